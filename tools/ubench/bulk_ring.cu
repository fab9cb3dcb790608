// Micro-benchmark: steady-state 1-D TMA ring as the fused kernel's mover warps run it, without any compute.
// NL lanes (spread over NW warps) each own one sub-slot per ring slot: wait load(k) -> store it (S2G) ->
// wait_group.read<1> -> load(k+D).  Reports bytes per clock per SM, in + out.
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
constexpr int kSlots = 4;
// mode: 0 = load+store, 1 = load only, 2 = store only
__global__ void k(const char* src, char* dst, int bytes, int lanes_per_warp, int nl, int items, long long item_stride, long long cta_stride,
                  int mode, long long* out) {
  extern __shared__ __align__(128) char sm[];
  __shared__ uint64_t bar[kSlots * 32];
  uint64_t pol_first, pol_last;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  for (int i = threadIdx.x; i < kSlots * 32; i += blockDim.x) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int me = warp * lanes_per_warp + lane;  // sub-slot
  if (lane >= lanes_per_warp || me >= nl) return;
  const int stride_s = (bytes + 127) & ~127;
  const char* g = src + (long long)blockIdx.x * cta_stride + (long long)me * bytes;
  char* gd = dst + (long long)blockIdx.x * cta_stride + (long long)me * bytes;
  auto sl = [&](int slot) { return smem_u32(sm + (slot * nl + me) * stride_s); };
  auto ld = [&](int item) {
    const int slot = item % kSlots;
    const uint32_t b = smem_u32(&bar[slot * 32 + me]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(sl(slot)),
                 "l"(g + (long long)item * item_stride), "r"(bytes), "r"(b), "l"(pol_first) : "memory");
  };
  const long long t0 = clock64();
  if (mode != 2) for (int i = 0; i < 3 && i < items; ++i) ld(i);
  for (int it = 0; it < items; ++it) {
    const int slot = it % kSlots;
    if (mode != 2) mbar_wait(smem_u32(&bar[slot * 32 + me]), (it / kSlots) & 1);
    if (mode != 1) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gd + (long long)it * item_stride),
                   "r"(sl(slot)), "r"(bytes), "l"(pol_last) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
    if (mode != 2 && it + 3 < items) ld(it + 3);
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  const long long t1 = clock64();
  if (me == 0) out[blockIdx.x] = t1 - t0;
}
int main() {
  char *src, *dst; long long* out;
  const int items = 256;
  cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const size_t cap = 148ull * items * 8 * 5024 + (1 << 20);
  cudaMalloc(&src, cap); cudaMalloc(&dst, cap);
  cudaMemset(src, 1, cap); cudaMemset(dst, 0, cap);
  struct Case { int bytes, lpw, nl; } cases[] = {{2512, 4, 8}, {2512, 1, 8}, {2512, 8, 8}, {5024, 4, 8}, {5024, 4, 4}, {1264, 4, 8}, {1264, 8, 16}, {10048, 4, 4}, {2512, 8, 16}};
  const char* mn[] = {"load+store", "load only ", "store only"};
  for (int resident = 1; resident >= 0; --resident)
    for (auto c : cases)
      for (int mode = 0; mode < 3; ++mode) {
        // resident: every item of a CTA maps to the same few addresses (L2 hits); else a fresh region per item (HBM)
        const long long item_stride = resident ? 0 : (long long)c.nl * c.bytes;
        const long long cta_stride = (long long)items * c.nl * c.bytes;
        const int nw = (c.nl + c.lpw - 1) / c.lpw;
        float ms = 0;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          k<<<148, 32 * nw, kSlots * c.nl * ((c.bytes + 127) & ~127)>>>(src, dst, c.bytes, c.lpw, c.nl, items, item_stride, cta_stride, mode, out);
          cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        long long h[148]; cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        const double moved = (double)c.bytes * c.nl * items * (mode == 0 ? 2 : 1);
        cudaError_t e = cudaGetLastError();
        printf("%s %s bytes=%5d lanes=%2d (per warp %d): %6.1f cycles per copy-slot, %6.2f B/clk/SM, %7.1f GB/s total  %s\n", resident ? "L2 " : "HBM", mn[mode],
               c.bytes, c.nl, c.lpw, avg / ((double)items * c.nl * (mode == 0 ? 2 : 1)), moved / avg, moved * 148 / (ms * 1e6), e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
