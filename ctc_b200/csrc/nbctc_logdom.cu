// Log-domain repair kernel of the single-label variant.
//
// The fast kernels (seqwarp_kernel.cuh, seqwide_kernel.cuh, stream_kernel.cuh) run the lattice in the LINEAR domain
// with emissions p_t(s) = softmax(x_t)[label_s] held in float32 and floored at 2^-120: a label whose logit lies more
// than 83 nats below its row's maximum cannot be represented, and a sequence whose admissible paths must pass through
// such an emission would get a wrong loss and gradient.  Every fast kernel therefore raises flag[b] when it floors an
// emission of a live state of sequence b, and this kernel -- launched behind it on the same stream with programmatic
// stream serialization; a small grid scans the flags, 8 per warp and step, and a warp that finds one set takes that
// sequence -- redoes those sequences in the LOG domain, where a log-probability
// of any size is just a number: float64 log alpha / log beta with a float32 correction term (logaddexp64, as the
// generic path, NoBlankCTC.py:16-19 _logsumexp), alpha checkpoints every 4 steps in the workspace record of the
// sequence, alpha replayed inside the tile next to beta, gamma = exp(log alpha + log beta - log Z).  It overwrites the
// loss and every gradient row of the sequence.  Slow (a double-precision exp per state and step) and rare.
#include <algorithm>

#include "common.cuh"

namespace nbctc {
namespace {

constexpr int kLogWarps = 4;
constexpr int kScan = 8;  // flags looked at by one warp per step
constexpr int kTT = 4;
constexpr unsigned kFull = 0xffffffffu;

template <int NS>
__device__ void logdom_sequence(const Problem& p, const LogWs& w, int64_t b, int lane, int* labs);

// A small grid scans the flags (the usual call has none set: the kernel is a few microseconds of flag reads); a warp
// that finds one redoes that sequence.
template <int NS>
__global__ void __launch_bounds__(kLogWarps * 32) logdom_kernel(Problem p, LogWs w) {
  constexpr int Lpad = 32 * NS;
  __shared__ int s_lab[kLogWarps][Lpad];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the fast kernel in front has finished and its stores are visible
  const int64_t nw = (int64_t)gridDim.x * kLogWarps;
  for (int64_t b0 = ((int64_t)blockIdx.x * kLogWarps + warp) * kScan; b0 < p.B; b0 += nw * kScan) {
    // kScan sequences per warp and step (few, so that many flagged sequences spread over many warps)
    const int64_t bl = b0 + lane;
    unsigned m = __ballot_sync(kFull, lane < kScan && bl < p.B && w.flag[bl] != 0);
    while (m) {
      const int i = __ffs(m) - 1;
      m &= m - 1;
      logdom_sequence<NS>(p, w, b0 + i, lane, s_lab[warp]);
      __syncwarp();
    }
  }
}

template <int NS>
__device__ void logdom_sequence(const Problem& p, const LogWs& w, int64_t b, int lane, int* labs) {
  constexpr int Lpad = 32 * NS;
  const int T = (int)p.T, B = (int)p.B, C = (int)p.C;
  const int Tb = (int)p.in_len[b], Lb = (int)p.tgt_len[b];  // flagged sequences are inside the parity domain
  const float wgt = p.w_scalar * (p.seq_w ? p.seq_w[b] : 1.f);
  int lab[NS], rank[NS];
  bool act[NS];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int s = lane * NS + j;
    act[j] = s < Lb;
    lab[j] = act[j] ? p.labels[b * p.Lmax + s] : 0;
    labs[s] = act[j] ? lab[j] : -1 - s;
  }
  __syncwarp();
  int R = 0;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int s = lane * NS + j;
    rank[j] = 0;
    if (act[j])
      for (int s2 = 0; s2 < s; ++s2) rank[j] += labs[s2] == lab[j];
    R = max(R, rank[j]);
  }
  R = __reduce_max_sync(kFull, R);
  const int64_t strideT = (int64_t)B * C;
  const float* x0 = p.logits + b * C;
  float* lse = w.lse_base + b * w.lse_stride;
  double* ckx = w.ckx_base + b * w.ckx_stride;

  // row log-partition (float32 max / sum, as the fast kernels) of row t
  auto row_lse = [&](const float* x) {
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, x[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(x[c] - m);
    s = warp_sum(s);
    return m + logf(s);
  };
  // log p_t(s) of the lane's states: exact for any gap
  auto log_emis = [&](const float* x, float l, double (&lp)[NS]) {
#pragma unroll
    for (int j = 0; j < NS; ++j) lp[j] = act[j] ? (double)x[lab[j]] - (double)l : -INFINITY;
  };
  auto alpha_step = [&](double (&a)[NS], const double (&lp)[NS]) {
    double up = __shfl_up_sync(kFull, a[NS - 1], 1);
    if (lane == 0) up = -INFINITY;
#pragma unroll
    for (int j = NS - 1; j >= 1; --j) a[j] = logaddexp64(a[j], a[j - 1]) + lp[j];
    a[0] = logaddexp64(a[0], up) + lp[0];
  };
  auto alpha_init = [&](double (&a)[NS], const double (&lp)[NS]) {
#pragma unroll
    for (int j = 0; j < NS; ++j) a[j] = (lane == 0 && j == 0) ? lp[0] : -INFINITY;
  };

  // ================================================================ phase 1
  double a[NS];
  for (int t = 0; t < Tb; ++t) {
    const float* x = x0 + (int64_t)t * strideT;
    const float l = row_lse(x);
    if (lane == 0) lse[t] = l;
    double lp[NS];
    log_emis(x, l, lp);
    if (t == 0) {
      alpha_init(a, lp);
    } else {
      if ((t & (kTT - 1)) == 0) {
#pragma unroll
        for (int j = 0; j < NS; ++j) ckx[(int64_t)(t / kTT) * Lpad + lane * NS + j] = a[j];
      }
      alpha_step(a, lp);
    }
  }
  const int sl = Lb - 1;
  double mine = a[0];
#pragma unroll
  for (int j = 1; j < NS; ++j) mine = (sl % NS >= j) ? a[j] : mine;
  const double ll = __shfl_sync(kFull, mine, sl / NS);
  if (lane == 0) p.loss[b] = (float)(-ll);
  if (p.grad == nullptr) return;
  float* g0 = p.grad + b * C;
  if (!(ll > -INFINITY) || wgt == 0.f) {  // (cannot happen inside the parity domain) / zero weight
    for (int t = 0; t < T; ++t)
      for (int c = lane; c < C; c += 32) g0[(int64_t)t * strideT + c] = 0.f;
    return;
  }
  __syncwarp();

  // ================================================================ phase 2
  double u[NS];  // log(beta_{t+1}(s) p_{t+1}(s))
  const int K = (Tb + kTT - 1) / kTT;
  for (int k = K - 1; k >= 0; --k) {
    const int t0 = k * kTT, nrow = min(kTT, Tb - t0);
    double at[kTT][NS], lpt[kTT][NS];
    double ar[NS];
    if (k > 0) {
#pragma unroll
      for (int j = 0; j < NS; ++j) ar[j] = ckx[(int64_t)k * Lpad + lane * NS + j];
    }
#pragma unroll
    for (int i = 0; i < kTT; ++i) {
      if (i < nrow) {
        const float* x = x0 + (int64_t)(t0 + i) * strideT;
        log_emis(x, lse[t0 + i], lpt[i]);
        if (t0 + i == 0) alpha_init(ar, lpt[i]);
        else alpha_step(ar, lpt[i]);
      }
#pragma unroll
      for (int j = 0; j < NS; ++j) at[i][j] = ar[j];
    }
#pragma unroll
    for (int i = kTT - 1; i >= 0; --i) {
      if (i < nrow) {
        const int t = t0 + i;
        double bt[NS];
        if (t == Tb - 1) {
#pragma unroll
          for (int j = 0; j < NS; ++j) bt[j] = (lane * NS + j == Lb - 1) ? 0.0 : -INFINITY;
        } else {
          double dn = __shfl_down_sync(kFull, u[0], 1);
          if (lane == 31) dn = -INFINITY;
#pragma unroll
          for (int j = 0; j < NS - 1; ++j) bt[j] = logaddexp64(u[j], u[j + 1]);
          bt[NS - 1] = logaddexp64(u[NS - 1], dn);
        }
        float gam[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          const double e = at[i][j] + bt[j] - ll;
          gam[j] = (act[j] && e > -745.0) ? (float)exp(e) : 0.f;
          u[j] = bt[j] + lpt[i][j];
        }
        const float* x = x0 + (int64_t)t * strideT;
        float* g = g0 + (int64_t)t * strideT;
        const float l = lse[t];
        for (int c = lane; c < C; c += 32) g[c] = wgt * expf(x[c] - l);
        __syncwarp();
        for (int r = 0; r <= R; ++r) {  // one state per label and round: plain read-modify-writes, ascending state order
#pragma unroll
          for (int j = 0; j < NS; ++j)
            if (act[j] && rank[j] == r) g[lab[j]] -= wgt * gam[j];
          __syncwarp();
        }
      }
    }
  }
  for (int t = Tb; t < T; ++t)
    for (int c = lane; c < C; c += 32) g0[(int64_t)t * strideT + c] = 0.f;
}

}  // namespace

int logdom_repair_launch(const Problem& p, const LogWs& w, cudaStream_t stream) {
  const unsigned grid = (unsigned)std::min<int64_t>((p.B + kLogWarps * kScan - 1) / (kLogWarps * kScan), 8 * 148);
  // programmatic stream serialization: the launch overlaps the tail of the fast kernel in front (which signals
  // launch_dependents where it supports it; behind any other kernel this is an ordinary launch)
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kLogWarps * 32);
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e;
  if (p.Lmax <= 32) e = cudaLaunchKernelEx(&cfg, logdom_kernel<1>, p, w);
  else if (p.Lmax <= 64) e = cudaLaunchKernelEx(&cfg, logdom_kernel<2>, p, w);
  else if (p.Lmax <= 128) e = cudaLaunchKernelEx(&cfg, logdom_kernel<4>, p, w);
  else if (p.Lmax <= 256) e = cudaLaunchKernelEx(&cfg, logdom_kernel<8>, p, w);
  else {
    set_error("log-domain repair kernel supports Lmax <= 256");
    return NBCTC_ERR_UNSUPPORTED;
  }
  NBCTC_CUDA_CHECK(e);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

}  // namespace nbctc
