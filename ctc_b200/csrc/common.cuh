// Shared helpers for the nbctc CUDA translation units (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "nbctc.h"

namespace nbctc {

// ---- error plumbing (thread-local message; nothing crosses the ABI but an int) -------
void set_error(const char* fmt, ...);
void clear_error();
extern std::atomic<uint64_t> g_launch_count;

#define NBCTC_CUDA_CHECK(expr)                                                          \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::nbctc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NBCTC_ERR_CUDA;                                                            \
    }                                                                                   \
  } while (0)

#define NBCTC_LAUNCH_CHECK()                                                            \
  do {                                                                                  \
    ::nbctc::g_launch_count.fetch_add(1, std::memory_order_relaxed);                    \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      ::nbctc::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NBCTC_ERR_CUDA;                                                            \
    }                                                                                   \
  } while (0)

// ---- problem description handed to the kernels ---------------------------------------
struct Problem {
  const float* logits;       // (T,B,C)
  const int32_t* labels;     // (B,Lmax)            [single-label variant]
  const float* targets;      // (B,Lmax,C)          [binary variant]
  const int64_t* in_len;     // (B)
  const int64_t* tgt_len;    // (B)
  float* loss;               // (B)
  double* loss_sum;          // scalar or null: sum_b loss[b] (float64, fixed order)
  float* loss_reduced;       // scalar or null: w_scalar * sum_b seq_w[b] * loss[b]
  float* grad;               // (T,B,C) or null
  const float* seq_w;        // (B) or null
  float w_scalar;
  bool sum_weighted;         // loss_sum = w_scalar * sum_b seq_w[b] * loss[b] (float64) instead of the plain sum
  int64_t T, B, C, Lmax;
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// generic (unfused) path
size_t generic_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax);
// gate: null, or a device flag -- the kernels run only if *gate != 0 (fallback of the tiled multi-label path)
int generic_launch(const Problem& p, bool binary, void* ws, size_t ws_bytes, cudaStream_t stream, const int* gate = nullptr);

// tiled path of the multi-label variant (nbctc_bin.cu): exact {0,1} targets with <= 31 classes per state; anything
// else is detected on the device and handed to the gated generic kernels through *flag_out
bool tiled_bin_supported(int64_t T, int64_t B, int64_t C, int64_t Lmax);
size_t tiled_bin_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax);
int tiled_bin_launch(const Problem& p, void* ws, size_t ws_bytes, cudaStream_t stream, const int** flag_out);

// fused path (single block-streaming kernel: row stream -> alpha, beta -> gradient; nbctc_stream.cu)
bool fused_supported(int64_t T, int64_t B, int64_t C, int64_t Lmax, bool binary);
size_t fused_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax, bool binary);
int fused_launch(const Problem& p, bool binary, void* ws, size_t ws_bytes, cudaStream_t stream);
bool fused_pointers_ok(const Problem& p);  // logits / grad 16-byte aligned (the bulk copies need it)
extern long long* g_stream_prof;           // role-profiler buffer (-DNBCTC_PROF builds), else null

// sequence-per-warp path of the single-label variant (nbctc_seqwarp.cu: one warp per sequence, whole batch in flight)
bool seqwarp_supported(int64_t T, int64_t B, int64_t C, int64_t Lmax);
bool seqwarp_is_wide(int64_t T, int64_t B, int64_t C, int64_t Lmax);  // rows through a shared-memory ring (needs 16-byte alignment)
size_t seqwarp_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax);
int seqwarp_launch(const Problem& p, void* ws, size_t ws_bytes, const float* row_lse_in, float* row_lse_out, cudaStream_t stream);

// log-domain repair of the sequences whose emissions the fast single-label kernels had to floor (nbctc_logdom.cu):
// flag[b] != 0 -> the sequence is redone; lse / ckx = scratch of the sequence (row log-partitions [T] float32, alpha
// checkpoints [ceil(T/4)][Lpad] float64, Lpad = 32/64/128/256 by Lmax), strides in elements
struct LogWs {
  const int* flag;       // [B] non-zero = redo the sequence
  float* lse_base;
  int64_t lse_stride;
  double* ckx_base;
  int64_t ckx_stride;
};
int logdom_repair_launch(const Problem& p, const LogWs& w, cudaStream_t stream);

// auxiliary cross-entropy on one frame per sequence, added into the gradient rows (nbctc_auxce.cu)
int aux_ce_launch(const float* logits, int64_t T, int64_t B, int64_t C, const int64_t* frame_index, const int64_t* in_len,
                  const int32_t* y_index, const float* y_multihot, int mode, float alpha_w, const float* seq_w, float* ce_per_seq,
                  float* grad, cudaStream_t stream);

// deterministic reduction of the per-sequence losses (float64, fixed order)
int reduce_loss_launch(const Problem& p, cudaStream_t stream);

#ifdef __CUDACC__
// ---- device helpers ------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// log(exp(a)+exp(b)) with a float64 state, -inf safe.  The correction term log1p(exp(-|a-b|)) is in [0, ln 2] and
// is evaluated in float32 (accurate expf/log1pf): its absolute error (~6e-8) does not grow with the state
// (float64 state + float32 correction keeps the gradient within 1e-7 at T = 4096).
__device__ __forceinline__ double logaddexp64(double a, double b) {
  double m = fmax(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + (double)log1pf(expf((float)(-fabs(a - b))));
}
// sequence is inside the parity domain (include/nbctc.h)
__device__ __forceinline__ bool seq_feasible(int64_t Tb, int64_t Lb, int64_t T, int64_t Lmax) {
  return Lb >= 1 && Lb <= Lmax && Tb >= Lb && Tb <= T;
}
#endif

}  // namespace nbctc
