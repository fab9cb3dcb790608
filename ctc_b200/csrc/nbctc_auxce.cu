// Auxiliary cross-entropy on one frame per sequence: the reference trains with
// Loss = CTC + alpha * CE (opts.py:74 --alpha, main.py:42, train.py:353), the CE taken on the scores of one frame of
// every sequence (train.py:434 classifies v_output[temporal-1]).  Two variants:
//   mode 0  nn.CrossEntropyLoss on a class index (models/__init__.py:85):  ce_b = lse(x) - x[y_b]
//   mode 1  the reference's own CrossEntropy module on a multi-hot row (CrossEntropy.py:17-32): q = softmax(x) (:22),
//           ce_b = log sum_c exp(q_c) (:25) - sum_{n: target[b][n] == 1} q_n (:26-29)
// One warp per sequence.  The kernel runs as the tail of the loss call on the same stream and ADDS
// alpha_w * d ce_b / d x into the gradient row (t_b, b) the loss kernel has just written: B rows out of T*B are
// touched, so the mix costs 1/T of the loss's own traffic and no second pass over the logits.
#include "common.cuh"

namespace nbctc {
namespace {

constexpr int kCeWarps = 8;

__global__ void __launch_bounds__(kCeWarps * 32) aux_ce_kernel(const float* __restrict__ logits, int64_t T, int64_t B, int64_t C,
                                                               const int64_t* __restrict__ frame_index,
                                                               const int64_t* __restrict__ in_len, const int32_t* __restrict__ y_index,
                                                               const float* __restrict__ y_multihot, int mode, float alpha_w,
                                                               const float* __restrict__ seq_w, float* __restrict__ ce_per_seq,
                                                               float* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * kCeWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const int64_t t = frame_index ? frame_index[b] : in_len[b] - 1;
  bool ok = t >= 0 && t < T;
  int y = 0;
  if (mode == 0) {
    y = y_index[b];
    ok = ok && y >= 0 && y < C;
  }
  if (!ok) {  // outside the domain: no contribution to the gradient
    if (lane == 0) ce_per_seq[b] = INFINITY;
    return;
  }
  const float* x = logits + (t * B + b) * C;
  float m = -INFINITY;
  for (int64_t c = lane; c < C; c += 32) m = fmaxf(m, x[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int64_t c = lane; c < C; c += 32) s += __expf(x[c] - m);
  s = warp_sum(s);
  const float lse = m + __logf(s);
  const float w = alpha_w * (seq_w ? seq_w[b] : 1.f);
  float* g = grad ? grad + (t * B + b) * C : nullptr;
  if (mode == 0) {
    if (lane == 0) ce_per_seq[b] = lse - x[y];
    if (g)
      for (int64_t c = lane; c < C; c += 32) g[c] += w * (__expf(x[c] - lse) - (c == y ? 1.f : 0.f));
    return;
  }
  // mode 1: q = softmax(x); S = sum exp(q); ce = log S - sum_pos q; d ce / d q_c = exp(q_c)/S - pos_c, then through
  // the softmax: d ce / d x_j = q_j (g_j - sum_c g_c q_c)
  const float* yr = y_multihot + b * C;
  float S = 0.f, pq = 0.f;
  for (int64_t c = lane; c < C; c += 32) {
    const float q = __expf(x[c] - lse);
    S += __expf(q);
    if (yr[c] == 1.f) pq += q;
  }
  S = warp_sum(S);
  pq = warp_sum(pq);
  if (lane == 0) ce_per_seq[b] = __logf(S) - pq;
  if (!g) return;
  float G = 0.f;
  for (int64_t c = lane; c < C; c += 32) {
    const float q = __expf(x[c] - lse);
    G += (__expf(q) / S - (yr[c] == 1.f ? 1.f : 0.f)) * q;
  }
  G = warp_sum(G);
  for (int64_t c = lane; c < C; c += 32) {
    const float q = __expf(x[c] - lse);
    g[c] += w * q * (__expf(q) / S - (yr[c] == 1.f ? 1.f : 0.f) - G);
  }
}

}  // namespace

int aux_ce_launch(const float* logits, int64_t T, int64_t B, int64_t C, const int64_t* frame_index, const int64_t* in_len,
                  const int32_t* y_index, const float* y_multihot, int mode, float alpha_w, const float* seq_w, float* ce_per_seq,
                  float* grad, cudaStream_t stream) {
  const unsigned blocks = (unsigned)((B + kCeWarps - 1) / kCeWarps);
  aux_ce_kernel<<<blocks, kCeWarps * 32, 0, stream>>>(logits, T, B, C, frame_index, in_len, y_index, y_multihot, mode, alpha_w, seq_w,
                                                      ce_per_seq, grad);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

}  // namespace nbctc
