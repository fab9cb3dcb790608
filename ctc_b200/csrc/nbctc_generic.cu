// Generic (unfused) path: three kernels with HBM intermediates.  Any T, C and Lmax <= 8192.
// This is the cross-check / fallback for shapes the fused kernel does not take; it is NOT
// the roofline path (it moves >= 12*T*B*C bytes plus the lattice tiles).
//
//   K1 rowstats : one warp per (t,b) row -> row log-partition + per-state raw emission
//                 (NoBlankCTC.py:136 log_softmax, :96-102 gather; binary: NoBlankBinaryCTC.py:109-112)
//   K2 lattice  : one CTA per sequence, thread = state, float64 log domain alpha then beta,
//                 gamma overwrites the emission tile (NoBlankCTC.py:71-87 transition, :58-68 read-out)
//   K3 grad     : one warp per (t,b) row -> w*(softmax - scatter(gamma)) / w*(sigmoid - gamma.y)/C
#include <algorithm>

#include "common.cuh"

namespace nbctc {

namespace {

constexpr int kRowWarps = 8;  // warps per CTA in the row kernels

struct GenericWs {
  float* rowc;    // (T,B)  row constant: lse (single) / mean softplus (binary)
  float* emis;    // (T,B,Lmax) raw emission, overwritten by gamma
  double* alpha;  // (T,B,Lmax)
  int* bad;       // (B) label-out-of-range flag
  int* rank;      // (B,Lmax) single-label: number of earlier states with the same label (scatter rounds of the gradient)
  int* maxrank;   // (B)
  const int* gate;  // null, or: run only if *gate != 0 (the tiled multi-label path declined the call, nbctc_bin.cu)
};

__host__ size_t carve(GenericWs* w, void* base, int64_t T, int64_t B, int64_t Lmax) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  size_t o_bad = take(sizeof(int) * B);
  size_t o_rank = take(sizeof(int) * B * Lmax);
  size_t o_maxrank = take(sizeof(int) * B);
  size_t o_rowc = take(sizeof(float) * T * B);
  size_t o_emis = take(sizeof(float) * T * B * Lmax);
  size_t o_alpha = take(sizeof(double) * T * B * Lmax);
  if (w) {
    char* c = static_cast<char*>(base);
    w->bad = reinterpret_cast<int*>(c + o_bad);
    w->rank = reinterpret_cast<int*>(c + o_rank);
    w->maxrank = reinterpret_cast<int*>(c + o_maxrank);
    w->rowc = reinterpret_cast<float*>(c + o_rowc);
    w->emis = reinterpret_cast<float*>(c + o_emis);
    w->alpha = reinterpret_cast<double*>(c + o_alpha);
  }
  return off;
}

// ------------------------------------------------------------------------------------
template <bool kBinary>
__global__ void __launch_bounds__(kRowWarps * 32)
rowstats_kernel(Problem p, GenericWs w) {
  if (w.gate != nullptr && *w.gate == 0) return;  // the tiled multi-label path took this call
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (row >= p.T * p.B) return;
  const int64_t t = row / p.B, b = row - t * p.B;
  const int64_t Tb = p.in_len[b], Lb = p.tgt_len[b];
  if (!seq_feasible(Tb, Lb, p.T, p.Lmax)) {
    if (t == 0 && lane == 0) w.bad[b] = 1;
    return;
  }
  const float* x = p.logits + row * p.C;
  if (t == 0) {  // validate labels once per sequence
    int bad = 0;
    if (!kBinary) {
      for (int64_t s = lane; s < Lb; s += 32) {
        int32_t l = p.labels[b * p.Lmax + s];
        bad |= (l < 0 || l >= p.C);
      }
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) w.bad[b] = bad;
  }
  if (t >= Tb) return;

  if (!kBinary) {
    float m = -INFINITY;
    for (int64_t c = lane; c < p.C; c += 32) m = fmaxf(m, x[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int64_t c = lane; c < p.C; c += 32) s += expf(x[c] - m);
    s = warp_sum(s);
    if (lane == 0) w.rowc[row] = m + logf(s);
    for (int64_t st = lane; st < Lb; st += 32) {
      int32_t l = p.labels[b * p.Lmax + st];
      float e = (l >= 0 && l < p.C) ? x[l] : 0.f;
      w.emis[row * p.Lmax + st] = e;
    }
  } else {
    // row constant (1/C) sum_c softplus(x_c); emission (1/C) y_s . x
    float sp = 0.f;
    for (int64_t c = lane; c < p.C; c += 32) {
      float v = x[c];
      sp += fmaxf(v, 0.f) + log1pf(expf(-fabsf(v)));
    }
    sp = warp_sum(sp);
    const float invC = 1.f / (float)p.C;
    if (lane == 0) w.rowc[row] = sp * invC;
    for (int64_t st = 0; st < Lb; ++st) {
      const float* y = p.targets + (b * p.Lmax + st) * p.C;
      float d = 0.f;
      for (int64_t c = lane; c < p.C; c += 32) d = fmaf(y[c], x[c], d);
      d = warp_sum(d);
      if (lane == 0) w.emis[row * p.Lmax + st] = d * invC;
    }
  }
}

// ------------------------------------------------------------------------------------
// The CTA-per-sequence kernels below walk their (virtual) block indices with a grid-stride loop: launched as the gated
// fallback of the tiled multi-label path they use a small grid, so that the usual "flag not set" case costs a few
// hundred CTAs that return at once instead of one per sequence and time chunk.
__device__ __forceinline__ void lattice_body(const Problem& p, const GenericWs& w, double* sm, int64_t b) {
  const int64_t Tb = p.in_len[b], Lb = p.tgt_len[b];
  const int64_t B = p.B, L = p.Lmax;
  if (!seq_feasible(Tb, Lb, p.T, L) || w.bad[b]) {
    if (threadIdx.x == 0) p.loss[b] = INFINITY;
    return;
  }
  double* a0 = sm;
  double* a1 = sm + L;
  const int nt = blockDim.x;
  if (p.labels != nullptr && p.grad != nullptr) {
    // duplicate rank of every state: the gradient kernel subtracts the gammas of one label in rank rounds, i.e. in
    // ascending state order without atomics (repeated labels accumulate; bit-reproducible)
    __shared__ int s_maxrank;
    if (threadIdx.x == 0) s_maxrank = 0;
    __syncthreads();
    const int32_t* lab = p.labels + b * L;
    for (int64_t s = threadIdx.x; s < Lb; s += nt) {
      const int32_t l = lab[s];
      int r = 0;
      for (int64_t s2 = 0; s2 < s; ++s2) r += lab[s2] == l;
      w.rank[b * L + s] = r;
      if (r) atomicMax(&s_maxrank, r);
    }
    __syncthreads();
    if (threadIdx.x == 0) w.maxrank[b] = s_maxrank;
  }
  // ---- alpha ----
  for (int64_t s = threadIdx.x; s < Lb; s += nt) {
    double v = (s == 0) ? (double)w.emis[(0 * B + b) * L] - (double)w.rowc[b] : -INFINITY;
    a0[s] = v;
    w.alpha[(0 * B + b) * L + s] = v;
  }
  __syncthreads();
  for (int64_t t = 1; t < Tb; ++t) {
    double* prev = (t & 1) ? a0 : a1;
    double* cur = (t & 1) ? a1 : a0;
    const int64_t row = t * B + b;
    const double rc = (double)w.rowc[row];
    for (int64_t s = threadIdx.x; s < Lb; s += nt) {
      double stay = prev[s];
      double adv = (s > 0) ? prev[s - 1] : -INFINITY;
      double v = logaddexp64(stay, adv) + ((double)w.emis[row * L + s] - rc);
      cur[s] = v;
      w.alpha[row * L + s] = v;
    }
    __syncthreads();
  }
  double* fin = ((Tb - 1) & 1) ? a1 : a0;
  const double ll = fin[Lb - 1];
  __syncthreads();
  if (threadIdx.x == 0) p.loss[b] = (float)(-ll);
  if (p.grad == nullptr) return;
  if (!(ll > -INFINITY)) return;  // cannot happen inside the parity domain
  // ---- beta / gamma (be = beta + emission, ping-pong in the same smem) ----
  {
    const int64_t row = (Tb - 1) * B + b;
    const double rc = (double)w.rowc[row];
    for (int64_t s = threadIdx.x; s < Lb; s += nt) {
      double beta = (s == Lb - 1) ? 0.0 : -INFINITY;
      double e = (double)w.emis[row * L + s] - rc;
      float g = expf((float)(w.alpha[row * L + s] + beta - ll));
      w.emis[row * L + s] = g;
      a0[s] = beta + e;
    }
    __syncthreads();
  }
  int flip = 0;
  for (int64_t t = Tb - 2; t >= 0; --t) {
    double* prev = flip ? a1 : a0;
    double* cur = flip ? a0 : a1;
    const int64_t row = t * B + b;
    const double rc = (double)w.rowc[row];
    for (int64_t s = threadIdx.x; s < Lb; s += nt) {
      double up = (s + 1 < Lb) ? prev[s + 1] : -INFINITY;
      double beta = logaddexp64(prev[s], up);
      double e = (double)w.emis[row * L + s] - rc;
      float g = expf((float)(w.alpha[row * L + s] + beta - ll));
      w.emis[row * L + s] = g;
      cur[s] = beta + e;
    }
    __syncthreads();
    flip ^= 1;
  }
}
template <bool kGated>
__global__ void __launch_bounds__(1024)
lattice_kernel(Problem p, GenericWs w) {
  extern __shared__ double sm[];  // [2][Lmax]
  if (!kGated) {
    lattice_body(p, w, sm, blockIdx.x);
    return;
  }
  if (*w.gate == 0) return;  // the tiled multi-label path took this call
  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    lattice_body(p, w, sm, b);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
template <bool kBinary>
__global__ void __launch_bounds__(kRowWarps * 32)
grad_kernel(Problem p, GenericWs w) {
  if (w.gate != nullptr && *w.gate == 0) return;  // the tiled multi-label path took this call
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (row >= p.T * p.B) return;
  const int64_t t = row / p.B, b = row - t * p.B;
  const int64_t Tb = p.in_len[b], Lb = p.tgt_len[b];
  float* g = p.grad + row * p.C;
  const float lossb = p.loss[b];
  if (!seq_feasible(Tb, Lb, p.T, p.Lmax) || t >= Tb || !(lossb < INFINITY)) {
    for (int64_t c = lane; c < p.C; c += 32) g[c] = 0.f;
    return;
  }
  const float* x = p.logits + row * p.C;
  float wgt = p.w_scalar * (p.seq_w ? p.seq_w[b] : 1.f);
  const float* gam = w.emis + row * p.Lmax;
  if (!kBinary) {
    const float lse = w.rowc[row];
    for (int64_t c = lane; c < p.C; c += 32) g[c] = wgt * expf(x[c] - lse);
    __syncwarp();
    const int mr = w.maxrank[b];
    for (int r = 0; r <= mr; ++r) {  // one state per label and round: plain read-modify-writes, ascending state order
      for (int64_t s = lane; s < Lb; s += 32)
        if (w.rank[b * p.Lmax + s] == r) {
          const int32_t l = p.labels[b * p.Lmax + s];
          g[l] -= wgt * gam[s];
        }
      __syncwarp();
    }
  } else {
    wgt /= (float)p.C;
    for (int64_t c = lane; c < p.C; c += 32) {
      float acc = 1.f / (1.f + expf(-x[c]));
      const float* y = p.targets + b * p.Lmax * p.C + c;
      for (int64_t s = 0; s < Lb; ++s) acc = fmaf(-gam[s], y[s * p.C], acc);
      g[c] = wgt * acc;
    }
  }
}

// ------------------------------------------------------------------------------------
// Binary variant with the sequence's multi-hot rows cached in shared memory.  One CTA = one sequence and a chunk
// of kBinTCh time steps; y[b] (L_b x C) is read from HBM once per CTA instead of once per (t,b) row.
// ys[s*Cp + c] with Cp odd: lanes that walk s (emissions) and lanes that walk c (gradient) are both conflict-free.
// When every entry of y[b] is exactly 0 or 1 (multi-hot, the reference's dataset: charades_ctc_pred.py:538-559) the
// dot products run over bit masks of the non-zero entries instead (state -> classes for the emissions, class ->
// states for the gradient); soft targets take the dense loops.
constexpr int kBinTCh = 128;

struct BinSmem {
  float* ys;        // [Lb][Cp]
  unsigned* smask;  // [Lb][CW]  classes of state s
  unsigned* cmask;  // [C][LW]   states that contain class c
  int* soft;        // != 0: some entry is neither 0 nor 1
  float* scratch;   // per-warp row buffer
};
__host__ __device__ inline size_t bin_smem_floats(int64_t Lmax, int64_t C, int Cp, int64_t per_warp) {
  const int64_t CW = (C + 31) / 32, LW = (Lmax + 31) / 32;
  return (size_t)(Lmax * Cp + Lmax * CW + C * LW + 4 + kRowWarps * per_warp);
}
__device__ __forceinline__ BinSmem bin_setup(const Problem& p, int64_t b, int Lb, int Cp, float* smf, int64_t per_warp) {
  const int C = (int)p.C, CW = (C + 31) / 32, LW = (Lb + 31) / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  BinSmem S;
  S.ys = smf;
  S.smask = reinterpret_cast<unsigned*>(smf + (size_t)p.Lmax * Cp);
  S.cmask = S.smask + (size_t)p.Lmax * CW;
  S.soft = reinterpret_cast<int*>(S.cmask + (size_t)C * ((p.Lmax + 31) / 32));
  S.scratch = reinterpret_cast<float*>(S.soft + 4) + (size_t)warp * per_warp;
  if (threadIdx.x == 0) *S.soft = 0;
  const float* y = p.targets + b * p.Lmax * p.C;
  for (int64_t i = threadIdx.x; i < (int64_t)Lb * C; i += blockDim.x) {
    const int s = (int)(i / C), c = (int)(i - (int64_t)s * C);
    S.ys[s * Cp + c] = y[i];
  }
  __syncthreads();
  int soft = 0;
  for (int i = warp; i < Lb * CW; i += kRowWarps) {  // state -> classes
    const int s = i / CW, wd = i - s * CW, c = wd * 32 + lane;
    const float v = c < C ? S.ys[s * Cp + c] : 0.f;
    soft |= (v != 0.f && v != 1.f);
    const unsigned m = __ballot_sync(0xffffffffu, v != 0.f);
    if (lane == 0) S.smask[s * CW + wd] = m;
  }
  for (int i = warp; i < C * LW; i += kRowWarps) {   // class -> states
    const int c = i / LW, wd = i - c * LW, st = wd * 32 + lane;
    const unsigned m = __ballot_sync(0xffffffffu, st < Lb && S.ys[st * Cp + c] != 0.f);
    if (lane == 0) S.cmask[c * LW + wd] = m;
  }
  if (soft) atomicOr(S.soft, 1);
  __syncthreads();
  return S;
}

// emissions e[t,b,s] = (1/C) y_s . x_t and the row constant (1/C) sum_c softplus(x_c) (NoBlankBinaryCTC.py:109-112)
__device__ __forceinline__ void rowstats_bin_smem_body(const Problem& p, const GenericWs& w, int Cp, float* smf, int64_t b,
                                                       int64_t chunk) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t Tb = p.in_len[b], Lb64 = p.tgt_len[b];
  if (!seq_feasible(Tb, Lb64, p.T, p.Lmax)) {
    if (chunk == 0 && threadIdx.x == 0) w.bad[b] = 1;
    return;
  }
  if (chunk == 0 && threadIdx.x == 0) w.bad[b] = 0;
  const int64_t t0 = chunk * kBinTCh;
  if (t0 >= Tb) return;
  const int Lb = (int)Lb64, C = (int)p.C, CW = (C + 31) / 32;
  const BinSmem S = bin_setup(p, b, Lb, Cp, smf, C);
  float* xs = S.scratch;  // this warp's logits row
  const bool soft = *S.soft != 0;
  const float invC = 1.f / (float)C;
  for (int64_t t = t0 + warp; t < min(t0 + kBinTCh, Tb); t += kRowWarps) {
    const int64_t row = t * p.B + b;
    const float* x = p.logits + row * C;
    float sp = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = x[c];
      xs[c] = v;
      sp += fmaxf(v, 0.f) + log1pf(expf(-fabsf(v)));
    }
    sp = warp_sum(sp);
    if (lane == 0) w.rowc[row] = sp * invC;
    __syncwarp();
    for (int s0 = 0; s0 < Lb; s0 += 32) {
      const int st = s0 + lane;
      float d0 = 0.f, d1 = 0.f;
      if (soft) {
        const float* yr = S.ys + (size_t)min(st, Lb - 1) * Cp;
        int c = 0;
        for (; c + 1 < C; c += 2) {
          d0 = fmaf(yr[c], xs[c], d0);
          d1 = fmaf(yr[c + 1], xs[c + 1], d1);
        }
        if (c < C) d0 = fmaf(yr[c], xs[c], d0);
      } else if (st < Lb) {
        for (int wd = 0; wd < CW; ++wd) {
          unsigned m = S.smask[st * CW + wd];
          while (m) {
            d0 += xs[wd * 32 + __ffs(m) - 1];
            m &= m - 1;
          }
        }
      }
      if (st < Lb) w.emis[row * p.Lmax + st] = (d0 + d1) * invC;
    }
    __syncwarp();
  }
}

template <bool kGated>
__global__ void __launch_bounds__(kRowWarps * 32)
rowstats_bin_smem_kernel(Problem p, GenericWs w, int Cp, int64_t nchunk) {
  extern __shared__ float smf[];
  if (!kGated) {
    rowstats_bin_smem_body(p, w, Cp, smf, blockIdx.x, blockIdx.y);
    return;
  }
  if (*w.gate == 0) return;  // the tiled multi-label path took this call
  for (int64_t v = blockIdx.x; v < p.B * nchunk; v += gridDim.x) {
    rowstats_bin_smem_body(p, w, Cp, smf, v % p.B, v / p.B);
    __syncthreads();
  }
}

// grad[t,b,c] = w/C * (sigmoid(x) - sum_s gamma_t(s) y[b,s,c])
__device__ __forceinline__ void grad_bin_smem_body(const Problem& p, const GenericWs& w, int Cp, float* smf, int64_t b,
                                                   int64_t chunk) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t Tb = p.in_len[b], Lb64 = p.tgt_len[b];
  const int C = (int)p.C;
  const int64_t t0 = chunk * kBinTCh;
  const int64_t t1 = min(t0 + kBinTCh, p.T);
  const float lossb = p.loss[b];
  const bool ok = seq_feasible(Tb, Lb64, p.T, p.Lmax) && (lossb < INFINITY);
  const int64_t tlive = ok ? min(t1, Tb) : t0;  // rows [tlive, t1) are zeros
  for (int64_t t = max(t0, tlive) + warp; t < t1; t += kRowWarps) {
    float* g = p.grad + (t * p.B + b) * C;
    for (int c = lane; c < C; c += 32) g[c] = 0.f;
  }
  if (!ok || t0 >= Tb) return;
  const int Lb = (int)Lb64, LW = (Lb + 31) / 32;
  const BinSmem S = bin_setup(p, b, Lb, Cp, smf, p.Lmax);
  float* gs = S.scratch;  // this warp's gamma row
  const bool soft = *S.soft != 0;
  const float wgt = p.w_scalar * (p.seq_w ? p.seq_w[b] : 1.f) / (float)C;
  for (int64_t t = t0 + warp; t < tlive; t += kRowWarps) {
    const int64_t row = t * p.B + b;
    const float* x = p.logits + row * C;
    const float* gam = w.emis + row * p.Lmax;
    for (int s = lane; s < Lb; s += 32) gs[s] = gam[s];
    __syncwarp();
    float* g = p.grad + row * C;
    for (int c = lane; c < C; c += 32) {
      float a0 = 1.f / (1.f + expf(-x[c])), a1 = 0.f;
      if (soft) {
        int s = 0;
        for (; s + 1 < Lb; s += 2) {
          a0 = fmaf(-gs[s], S.ys[s * Cp + c], a0);
          a1 = fmaf(-gs[s + 1], S.ys[(s + 1) * Cp + c], a1);
        }
        if (s < Lb) a0 = fmaf(-gs[s], S.ys[s * Cp + c], a0);
      } else {
        for (int wd = 0; wd < LW; ++wd) {
          unsigned m = S.cmask[c * LW + wd];
          while (m) {
            a1 -= gs[wd * 32 + __ffs(m) - 1];
            m &= m - 1;
          }
        }
      }
      g[c] = wgt * (a0 + a1);
    }
    __syncwarp();
  }
}

template <bool kGated>
__global__ void __launch_bounds__(kRowWarps * 32)
grad_bin_smem_kernel(Problem p, GenericWs w, int Cp, int64_t nchunk) {
  extern __shared__ float smf[];
  if (!kGated) {
    grad_bin_smem_body(p, w, Cp, smf, blockIdx.x, blockIdx.y);
    return;
  }
  if (*w.gate == 0) return;  // the tiled multi-label path took this call
  for (int64_t v = blockIdx.x; v < p.B * nchunk; v += gridDim.x) {
    grad_bin_smem_body(p, w, Cp, smf, v % p.B, v / p.B);
    __syncthreads();
  }
}

__global__ void reduce_loss_kernel(const float* __restrict__ loss, const float* __restrict__ seq_w, float w_scalar,
                                   int64_t B, double* __restrict__ sum_out, float* __restrict__ reduced_out, bool sum_weighted) {
  // one CTA, fixed-order tree => bit-reproducible
  __shared__ double sh[2][32];
  double acc = 0.0, wacc = 0.0;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
    double l = (double)loss[i];
    acc += l;
    wacc += seq_w ? l * (double)seq_w[i] : l;
  }
  acc = warp_sum(acc);
  wacc = warp_sum(wacc);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = acc; sh[1][threadIdx.x >> 5] = wacc; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const bool in = threadIdx.x < (blockDim.x >> 5);
    double v = warp_sum(in ? sh[0][threadIdx.x] : 0.0);
    double wv = warp_sum(in ? sh[1][threadIdx.x] : 0.0);
    if (threadIdx.x == 0) {
      if (sum_out) *sum_out = sum_weighted ? wv * (double)w_scalar : v;
      if (reduced_out) *reduced_out = (float)(wv * (double)w_scalar);
    }
  }
}

}  // namespace

size_t generic_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax) {
  (void)C;
  return carve(nullptr, nullptr, T, B, Lmax);
}

int reduce_loss_launch(const Problem& p, cudaStream_t stream) {
  reduce_loss_kernel<<<1, 1024, 0, stream>>>(p.loss, p.seq_w, p.w_scalar, p.B, p.loss_sum, p.loss_reduced, p.sum_weighted);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

int generic_launch(const Problem& p, bool binary, void* ws, size_t ws_bytes, cudaStream_t stream, const int* gate) {
  if (p.Lmax > 8192) {
    set_error("generic path supports Lmax <= 8192 (got %lld)", (long long)p.Lmax);
    return NBCTC_ERR_UNSUPPORTED;
  }
  GenericWs w;
  w.gate = gate;
  size_t need = carve(&w, ws, p.T, p.B, p.Lmax);
  if (ws == nullptr || ws_bytes < need) {
    set_error("workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    return NBCTC_ERR_WORKSPACE;
  }
  const int64_t rows = p.T * p.B;
  const unsigned row_blocks = (unsigned)((rows + kRowWarps - 1) / kRowWarps);
  // binary variant: the sequence's multi-hot rows are cached in shared memory when they fit
  const int Cp = (int)(p.C | 1);
  const size_t bin_smem_rs = sizeof(float) * bin_smem_floats(p.Lmax, p.C, Cp, p.C);
  const size_t bin_smem_gr = sizeof(float) * bin_smem_floats(p.Lmax, p.C, Cp, p.Lmax);
  const bool bin_smem = binary && std::max(bin_smem_rs, bin_smem_gr) <= 200 * 1024;
  // gated fallback: a small grid (the kernels loop over their virtual blocks), else one CTA per sequence and chunk
  const int64_t nchunk = (p.T + kBinTCh - 1) / kBinTCh;
  const int64_t cap = 148 * 4;
  const dim3 bin_grid_full((unsigned)p.B, (unsigned)nchunk);
  const unsigned bin_grid_gated = (unsigned)std::min<int64_t>(p.B * nchunk, cap);
  if (bin_smem) {
    if (bin_smem_rs > 48 * 1024) {
      NBCTC_CUDA_CHECK(cudaFuncSetAttribute(rowstats_bin_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_smem_rs));
      NBCTC_CUDA_CHECK(cudaFuncSetAttribute(rowstats_bin_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_smem_rs));
    }
    if (gate) rowstats_bin_smem_kernel<true><<<bin_grid_gated, kRowWarps * 32, bin_smem_rs, stream>>>(p, w, Cp, nchunk);
    else rowstats_bin_smem_kernel<false><<<bin_grid_full, kRowWarps * 32, bin_smem_rs, stream>>>(p, w, Cp, nchunk);
  } else if (binary) {
    rowstats_kernel<true><<<row_blocks, kRowWarps * 32, 0, stream>>>(p, w);
  } else {
    rowstats_kernel<false><<<row_blocks, kRowWarps * 32, 0, stream>>>(p, w);
  }
  NBCTC_LAUNCH_CHECK();

  int nt = (int)std::min<int64_t>(1024, (p.Lmax + 31) / 32 * 32);
  size_t smem = 2 * sizeof(double) * p.Lmax;
  if (smem > 48 * 1024) {
    NBCTC_CUDA_CHECK(cudaFuncSetAttribute(lattice_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NBCTC_CUDA_CHECK(cudaFuncSetAttribute(lattice_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (gate) lattice_kernel<true><<<(unsigned)std::min<int64_t>(p.B, cap), nt, smem, stream>>>(p, w);
  else lattice_kernel<false><<<(unsigned)p.B, nt, smem, stream>>>(p, w);
  NBCTC_LAUNCH_CHECK();

  if (p.grad) {
    if (bin_smem) {
      if (bin_smem_gr > 48 * 1024) {
        NBCTC_CUDA_CHECK(cudaFuncSetAttribute(grad_bin_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_smem_gr));
        NBCTC_CUDA_CHECK(cudaFuncSetAttribute(grad_bin_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_smem_gr));
      }
      if (gate) grad_bin_smem_kernel<true><<<bin_grid_gated, kRowWarps * 32, bin_smem_gr, stream>>>(p, w, Cp, nchunk);
      else grad_bin_smem_kernel<false><<<bin_grid_full, kRowWarps * 32, bin_smem_gr, stream>>>(p, w, Cp, nchunk);
    } else if (binary)
      grad_kernel<true><<<row_blocks, kRowWarps * 32, 0, stream>>>(p, w);
    else
      grad_kernel<false><<<row_blocks, kRowWarps * 32, 0, stream>>>(p, w);
    NBCTC_LAUNCH_CHECK();
  }
  return NBCTC_OK;
}

}  // namespace nbctc
