// Micro-benchmark: dependent-chain latency and throughput of the FP64 pipe, SHFL and LDS on the target GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void lat(double* out, long long* cyc, int iters, double seed, float fseed) {
  double a = seed + threadIdx.x, b = seed * 0.5, c = 1.0000001;
  float f = fseed;
  __shared__ float sm[1024];
  sm[threadIdx.x] = fseed + threadIdx.x;
  __syncthreads();
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { a = a + b; a = a + b; a = a + b; a = a + b; }                 // DADD chain
    if (MODE == 1) { a = a * c; a = a * c; a = a * c; a = a * c; }                 // DMUL chain
    if (MODE == 2) { a = fma(a, c, b); a = fma(a, c, b); a = fma(a, c, b); a = fma(a, c, b); }
    if (MODE == 3) { a = __shfl_up_sync(0xffffffffu, a, 1, 16); a = __shfl_up_sync(0xffffffffu, a, 1, 16);
                     a = __shfl_up_sync(0xffffffffu, a, 1, 16); a = __shfl_up_sync(0xffffffffu, a, 1, 16); }
    if (MODE == 4) { f = (float)((double)f * c); f = (float)((double)f * c); f = (float)((double)f * c); f = (float)((double)f * c); }
    if (MODE == 5) { idx = (int)sm[idx & 1023]; idx = (int)sm[idx & 1023]; idx = (int)sm[idx & 1023]; idx = (int)sm[idx & 1023]; }
    if (MODE == 6) { f = f * 1.0000001f + 0.5f; f = f * 1.0000001f + 0.5f; f = f * 1.0000001f + 0.5f; f = f * 1.0000001f + 0.5f; }
    if (MODE == 7) { // the chain step: shfl + dadd + dmul
      double up = __shfl_up_sync(0xffffffffu, a, 1, 16); b = (b + a) * c; a = (a + up) * c;
      up = __shfl_up_sync(0xffffffffu, a, 1, 16); b = (b + a) * c; a = (a + up) * c;
      up = __shfl_up_sync(0xffffffffu, a, 1, 16); b = (b + a) * c; a = (a + up) * c;
      up = __shfl_up_sync(0xffffffffu, a, 1, 16); b = (b + a) * c; a = (a + up) * c; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + f + idx;
}
template <int MODE> void run(const char* name, int warps, int blocks) {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&cyc, 8 * 4096);
  int iters = 4096;
  lat<MODE><<<blocks, warps * 32>>>(out, cyc, iters, 1.0, 1.0f);
  lat<MODE><<<blocks, warps * 32>>>(out, cyc, iters, 1.0, 1.0f);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps/SM=%2d : %.1f cycles per op (per warp), %.2f warp-ops/cycle/SM\n", name, warps, (double)h / (iters * 4.0),
         warps * iters * 4.0 / (double)h);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 4, 8, 16, 32}) {
    run<0>("DADD dependent", w, 148);
    run<1>("DMUL dependent", w, 148);
    run<2>("DFMA dependent", w, 148);
  }
  run<3>("SHFL.64 dependent", 1, 148);
  run<4>("F2F f32->f64->DMUL->f32", 1, 148);
  run<5>("LDS dependent (+cvt)", 1, 148);
  run<6>("FFMA dependent", 1, 148);
  run<7>("chain step (shfl,2dadd,2dmul)", 1, 148);
  run<7>("chain step (shfl,2dadd,2dmul)", 4, 148);
  return 0;
}
