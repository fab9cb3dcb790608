// Micro-benchmark: cp.async.bulk (1-D TMA) global->shared and shared->global cost per copy as a function of size and
// of the 128-byte phase of the global and shared addresses.  One thread per CTA issues the copies, 148 CTAs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
// mode 0: G2S, wait for each batch of `nb` copies; mode 1: S2G with commit/wait_group.read per batch
__global__ void k(const char* src, char* dst, int bytes, int goff, int soff, int nb, int iters, int mode, long long gstride, long long* out, int hint) {
  uint64_t pol;
  if (hint == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  extern __shared__ __align__(128) char sm[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int stride_s = (bytes + 256 + 127) & ~127;
  const char* g = src + (long long)blockIdx.x * gstride * 64 + goff;
  char* gd = dst + (long long)blockIdx.x * gstride * 64 + goff;
  long long t_issue = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const long long a = clock64();
    if (mode == 0) {
      for (int i = 0; i < nb; ++i)
        if (hint)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(sm + i * stride_s + soff)),
                       "l"(g + ((long long)(it * nb + i) % 64) * gstride), "r"(bytes), "r"(smem_u32(&bar)), "l"(pol) : "memory");
        else
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + i * stride_s + soff)),
                       "l"(g + ((long long)(it * nb + i) % 64) * gstride), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes * nb) : "memory");
      t_issue += clock64() - a;
      mbar_wait(smem_u32(&bar), it & 1);
    } else {
      for (int i = 0; i < nb; ++i)
        if (hint)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gd + ((long long)(it * nb + i) % 64) * gstride),
                       "r"(smem_u32(sm + i * stride_s + soff)), "r"(bytes), "l"(pol) : "memory");
        else
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gd + ((long long)(it * nb + i) % 64) * gstride),
                       "r"(smem_u32(sm + i * stride_s + soff)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      t_issue += clock64() - a;
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  const long long t1 = clock64();
  if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue; }
}
int main() {
  char *src, *dst; long long* out;
  const long long gstride = 2572288;  // 4096*628: row stride of the cfg2 tensor
  cudaMalloc(&src, 148ll * 64 * gstride + 4096); cudaMalloc(&dst, 148ll * 64 * gstride + 4096); cudaMalloc(&out, 16);
  cudaMemset(src, 1, 148ll * 64 * gstride + 4096);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int nb = 8, iters = 200;
  struct Case { int bytes, goff, soff; } cases[] = {
      {2512, 0, 0}, {2512, 80, 0}, {640, 0, 0}, {10048, 0, 0}};
  for (int hint = 0; hint < 3; ++hint)
  for (int mode = 0; mode < 2; ++mode)
    for (auto c : cases) {
      const int nbb = c.bytes > 12000 ? 4 : nb;
      for (int rep = 0; rep < 2; ++rep) k<<<148, 32, 200 * 1024>>>(src, dst, c.bytes, c.goff, c.soff, nbb, iters, mode, gstride, out, hint);
      long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      printf("hint=%d %s bytes=%5d goff=%3d soff=%3d : %7.1f cycles/copy total (%6.1f issue), %6.2f B/clk/SM  %s\n", hint, mode ? "S2G" : "G2S", c.bytes, c.goff, c.soff,
             (double)h[0] / (iters * nbb), (double)h[1] / (iters * nbb), (double)c.bytes * iters * nbb / h[0], e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
