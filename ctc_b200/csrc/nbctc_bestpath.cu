// Max-plus (Viterbi) alignment on the no-blank lattice + per-frame argmax.
// Integer outputs are bit-exact against oracle/restatement.py::best_path because the scores
// are float64 sums of the raw fp32 logits evaluated in the same order (add and max only).
#include "common.cuh"

namespace nbctc {
namespace {

constexpr int kRowWarps = 8;

__global__ void __launch_bounds__(kRowWarps * 32)
frame_argmax_kernel(const float* __restrict__ x, int64_t rows, int64_t C, int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* r = x + row * C;
  float best = -INFINITY;
  int64_t bi = C;  // sentinel: nothing seen
  for (int64_t c = lane; c < C; c += 32) {
    float v = r[c];
    if (v > best || (bi == C && !(v != v))) { best = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o);
    int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) out[row] = (int32_t)(bi == C ? 0 : bi);
}

// one CTA per sequence; thread = state; back-pointers in global scratch (T,B,Lmax) bytes
__global__ void __launch_bounds__(1024)
best_path_kernel(const float* __restrict__ x, int64_t T, int64_t B, int64_t C,
                 const int32_t* __restrict__ labels, int64_t L,
                 const int64_t* __restrict__ in_len, const int64_t* __restrict__ tgt_len,
                 int32_t* __restrict__ states, double* __restrict__ score, uint8_t* __restrict__ back) {
  extern __shared__ double sm[];  // [2][L]
  const int64_t b = blockIdx.x;
  const int64_t Tb = in_len[b], Lb = tgt_len[b];
  const int nt = blockDim.x;
  for (int64_t t = threadIdx.x; t < T; t += nt) states[b * T + t] = -1;
  bool ok = seq_feasible(Tb, Lb, T, L);
  if (ok) {
    int bad = 0;
    for (int64_t s = threadIdx.x; s < Lb; s += nt) {
      int32_t l = labels[b * L + s];
      bad |= (l < 0 || l >= C);
    }
    ok = !__syncthreads_or(bad);
  }
  if (!ok) {
    if (threadIdx.x == 0 && score) score[b] = -INFINITY;
    return;
  }
  __syncthreads();
  double* v0 = sm;
  double* v1 = sm + L;
  for (int64_t s = threadIdx.x; s < Lb; s += nt)
    v0[s] = (s == 0) ? (double)x[(0 * B + b) * C + labels[b * L]] : -INFINITY;
  __syncthreads();
  for (int64_t t = 1; t < Tb; ++t) {
    double* prev = (t & 1) ? v0 : v1;
    double* cur = (t & 1) ? v1 : v0;
    const float* r = x + (t * B + b) * C;
    for (int64_t s = threadIdx.x; s < Lb; s += nt) {
      double stay = prev[s];
      double adv = (s > 0) ? prev[s - 1] : -INFINITY;
      bool take = adv > stay;  // tie -> stay
      cur[s] = (take ? adv : stay) + (double)r[labels[b * L + s]];
      back[(t * B + b) * L + s] = take ? 1 : 0;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double* fin = ((Tb - 1) & 1) ? v1 : v0;
    if (score) score[b] = fin[Lb - 1];
    int64_t s = Lb - 1;
    for (int64_t t = Tb - 1; t >= 0; --t) {
      states[b * T + t] = (int32_t)s;
      if (t > 0 && back[(t * B + b) * L + s]) --s;
    }
  }
}

}  // namespace
}  // namespace nbctc

using namespace nbctc;

extern "C" size_t nbctc_best_path_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax) {
  (void)C;
  if (T < 1 || B < 1 || Lmax < 1) return 0;
  return align_up((size_t)T * B * Lmax, 256);
}

extern "C" int nbctc_best_path_i32(const float* logits, int64_t T, int64_t B, int64_t C, const int32_t* labels,
                                   int64_t Lmax, const int64_t* input_lengths, const int64_t* target_lengths,
                                   int32_t* states, double* score, int32_t* argmax, void* workspace,
                                   size_t workspace_bytes, nbctc_stream_t stream) {
  clear_error();
  if (!logits || T < 1 || B < 1 || C < 1) {
    set_error("invalid argument to nbctc_best_path_i32");
    return NBCTC_ERR_INVALID_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (argmax) {
    const int64_t rows = T * B;
    frame_argmax_kernel<<<(unsigned)((rows + kRowWarps - 1) / kRowWarps), kRowWarps * 32, 0, st>>>(logits, rows, C, argmax);
    NBCTC_LAUNCH_CHECK();
  }
  if (states) {
    if (!labels || !input_lengths || !target_lengths || Lmax < 1) {
      set_error("labels/lengths required for the alignment");
      return NBCTC_ERR_INVALID_ARG;
    }
    if (Lmax > 8192) {
      set_error("best path supports Lmax <= 8192");
      return NBCTC_ERR_UNSUPPORTED;
    }
    size_t need = nbctc_best_path_workspace_bytes(T, B, C, Lmax);
    if (!workspace || workspace_bytes < need) {
      set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
      return NBCTC_ERR_WORKSPACE;
    }
    int nt = (int)std::min<int64_t>(1024, (Lmax + 31) / 32 * 32);
    size_t smem = 2 * sizeof(double) * Lmax;
    if (smem > 48 * 1024)
      NBCTC_CUDA_CHECK(cudaFuncSetAttribute(best_path_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    best_path_kernel<<<(unsigned)B, nt, smem, st>>>(logits, T, B, C, labels, Lmax, input_lengths, target_lengths,
                                                    states, score, static_cast<uint8_t*>(workspace));
    NBCTC_LAUNCH_CHECK();
  }
  return NBCTC_OK;
}
