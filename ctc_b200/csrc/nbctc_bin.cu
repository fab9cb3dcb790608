// Tiled path of the multi-label variant (NoBlankBinaryCTC.py): five launches, time-batched row kernels.
//
// The multi-hot targets y[b] (L_b x C) do not change over time, so every product with them is organised as
// "decode an index once, use it for kTB = 8 time steps":
//
//   P  bin_prepass   : one warp per (b,s) target row -> the state's class list (<= 31 byte indices) and the
//                      class -> states bit masks; raises a flag when a row is not exact {0,1} or holds more than
//                      31 classes -- every kernel below then returns at once and the generic kernels (gated on the
//                      same flag) run instead.
//   K1 bin_emis      : CTA = one sequence x 256 time steps, warp = batches of 8 rows staged in shared memory;
//                      row constant (1/C) sum_c softplus(x_c) and emissions e_t(s) = (1/C) sum_{c in S_s} x_t(c)
//                      with lane = state (NoBlankBinaryCTC.py:109-112: -BCELoss(sigmoid(x_t), y_s)).
//   K2 lattice_tile  : warp = sequence; the float64 linear-domain chain of the fused kernel (stream_kernel.cuh:
//                      16 lanes x NS states per direction, exact power-of-two rescaling, one alpha checkpoint per
//                      tile of 8 steps, alpha replay next to beta in phase 2) on the emission tiles;
//                      gamma overwrites the emissions (NoBlankBinaryCTC.py:72-95 transition, read-out :58-68).
//   K3 bin_grad      : CTA = one sequence x 256 time steps, warp = batches of 8 rows, lane = class:
//                      grad = w/C * (sigmoid(x) - sum_{s in M_c} gamma_t(s)), the states of a class walked over its
//                      bit mask once per batch (ascending state order: deterministic).
//
// HBM traffic: logits read twice, gradient written once, emission/gamma tile (T,B,Lmax) fp32 written twice and read
// twice, checkpoints (T/8,B,Lpad) f64 -- about 1.8x the algorithmic bytes at C = 157, Lmax = 32.
#include <algorithm>

#include "common.cuh"
#include "stream_kernel.cuh"

namespace nbctc {

namespace {

constexpr int kTB = 8;        // rows (time steps) per warp batch
constexpr int kRowWarps = 8;  // warps per CTA in K1 / K3
constexpr int kTCh = 256;     // time steps per CTA in K1 / K3
constexpr int kLatWarps = 4;  // sequences per CTA in K2

struct TiledWs {
  int* flag;            // != 0: the targets are outside this path's domain -> generic kernels
  uint32_t* lists;      // [B][Lmax][8]  byte 0 = class count (<= 31), bytes 1.. = class indices
  uint32_t* cmask;      // [B][C][LW]    states that contain the class
  double* ckpt;         // [B][NT][Lpad] alpha checkpoints
  int* cke;             // [B][NT]       their exponents
  float* rowc;          // (T,B)         (1/C) sum_c softplus(x_c)
  float* emis;          // (T,B,Lmax)    emissions, overwritten by gamma
  int LW, NT, Lpad;
};

struct Layout {
  size_t o_lists, o_cmask, o_ckpt, o_cke, o_rowc, o_emis, total;
  int LW, NT, Lpad;
};

Layout layout(int64_t T, int64_t B, int64_t C, int64_t Lmax) {
  Layout l{};
  l.Lpad = Lmax <= 32 ? 32 : Lmax <= 64 ? 64 : Lmax <= 128 ? 128 : 256;
  l.LW = (int)((Lmax + 31) / 32);
  l.NT = (int)((T + 7) / 8);
  size_t off = 256;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  l.o_lists = take((size_t)B * Lmax * 32);
  l.o_cmask = take(sizeof(uint32_t) * (size_t)B * C * l.LW);
  l.o_ckpt = take(sizeof(double) * (size_t)B * l.NT * l.Lpad);
  l.o_cke = take(sizeof(int) * (size_t)B * l.NT);
  l.o_rowc = take(sizeof(float) * (size_t)T * B);
  l.o_emis = take(sizeof(float) * (size_t)T * B * Lmax);
  l.total = off;
  return l;
}

// ------------------------------------------------------------------------------------ P: class lists and masks
__global__ void __launch_bounds__(256) bin_prepass_kernel(Problem p, TiledWs w) {
  __shared__ uint32_t rec[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= p.B * p.Lmax) return;
  const int64_t b = row / p.Lmax, s = row - b * p.Lmax;
  const int64_t Tb = p.in_len[b], Lb = p.tgt_len[b];
  const bool valid = seq_feasible(Tb, Lb, p.T, p.Lmax) && s < Lb;
  if (lane < 8) rec[warp][lane] = 0u;
  __syncwarp();
  if (valid) {
    unsigned char* bytes = reinterpret_cast<unsigned char*>(rec[warp]);
    const float* y = p.targets + row * p.C;
    const int C = (int)p.C;
    int count = 0;
    bool bad = false;
    float yv[8];  // C <= 256: the whole row, loaded ahead of the ballots and atomics
#pragma unroll
    for (int i = 0; i < 8; ++i) yv[i] = lane + 32 * i < C ? __ldg(y + lane + 32 * i) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c0 = 32 * i;
      if (c0 >= C) break;
      const int c = c0 + lane;
      const float v = yv[i];
      const bool one = v == 1.f;
      bad |= !(one || v == 0.f);
      const unsigned m = __ballot_sync(0xffffffffu, one);
      const int pos = count + __popc(m & ((1u << lane) - 1u));
      if (one && pos < 31) bytes[1 + pos] = (unsigned char)c;
      if (one) atomicOr(&w.cmask[((size_t)b * C + c) * w.LW + (s >> 5)], 1u << (s & 31));
      count += __popc(m);
    }
    bad = __any_sync(0xffffffffu, bad) || count > 31;
    if (lane == 0) {
      bytes[0] = (unsigned char)min(count, 31);
      if (bad) atomicOr(w.flag, 1);
    }
    __syncwarp();
  }
  if (lane < 8) w.lists[row * 8 + lane] = rec[warp][lane];
}

__device__ __forceinline__ float softplus_fast(float v) {
  return fmaxf(v, 0.f) + __logf(1.f + __expf(-fabsf(v)));
}

// ------------------------------------------------------------------------------------ K1: emissions
template <int NCI>
__global__ void __launch_bounds__(kRowWarps * 32, NCI <= 5 ? 4 : 2) bin_emis_kernel(Problem p, TiledWs w) {
  if (*w.flag != 0) return;
  constexpr int Cp = 32 * NCI + 8;  // row stride: rows start 8 banks apart
  extern __shared__ float smf[];  // [kRowWarps][kTB][Cp]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = blockIdx.x;
  const int64_t Tb = p.in_len[b], Lb64 = p.tgt_len[b];
  if (!seq_feasible(Tb, Lb64, p.T, p.Lmax)) return;
  const int64_t t0 = (int64_t)blockIdx.y * kTCh;
  if (t0 >= Tb) return;
  const int64_t tend = min(t0 + kTCh, Tb);
  const int Lb = (int)Lb64, C = (int)p.C;
  const float invC = 1.f / (float)C;
  float* xs = smf + (size_t)warp * kTB * Cp;
  const int64_t rstride = p.B * p.C;  // floats between consecutive time steps of a sequence
  // the whole batch of a warp is in flight before its first use: one memory round trip per batch.  (Requesting the
  // next batch ahead of the walk keeps v live across it: 124 registers, two CTAs per SM, 0.34 instead of 0.26 ms.)
  float v[NCI][kTB];
  auto request = [&](int64_t tb0) {
    const int nrow = (int)min((int64_t)kTB, tend - tb0);
    const float* x0 = p.logits + (tb0 * p.B + b) * p.C;
#pragma unroll
    for (int i = 0; i < NCI; ++i) {
      const int c = lane + 32 * i;
#pragma unroll
      for (int r = 0; r < kTB; ++r) v[i][r] = (c < C && r < nrow) ? __ldg(x0 + r * rstride + c) : 0.f;
    }
  };
  for (int64_t tb0 = t0 + (int64_t)warp * kTB; tb0 < tend; tb0 += kRowWarps * kTB) {
    const int nrow = (int)min((int64_t)kTB, tend - tb0);
    request(tb0);
    float sp[kTB];
#pragma unroll
    for (int r = 0; r < kTB; ++r) sp[r] = 0.f;
#pragma unroll
    for (int i = 0; i < NCI; ++i) {
      const int c = lane + 32 * i;
#pragma unroll
      for (int r = 0; r < kTB; ++r) {
        xs[r * Cp + c] = v[i][r];
        sp[r] += c < C ? softplus_fast(v[i][r]) : 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < kTB; ++r) {
      sp[r] = warp_sum(sp[r]);
      if (lane == r && r < nrow) w.rowc[(tb0 + r) * p.B + b] = sp[r] * invC;
    }
    __syncwarp();
    for (int s0 = 0; s0 < Lb; s0 += 32) {
      const int st = s0 + lane;
      const bool valid = st < Lb;
      uint32_t lw[8];
      {
        const uint4* src = reinterpret_cast<const uint4*>(w.lists + ((size_t)b * p.Lmax + min(st, Lb - 1)) * 8);
        const uint4 a = __ldg(src), c4 = __ldg(src + 1);
        lw[0] = a.x; lw[1] = a.y; lw[2] = a.z; lw[3] = a.w;
        lw[4] = c4.x; lw[5] = c4.y; lw[6] = c4.z; lw[7] = c4.w;
      }
      const int n = valid ? (int)(lw[0] & 0xffu) : 0;
      const int words = (__reduce_max_sync(0xffffffffu, n) + 4) >> 2;  // bytes 0..n
      float d[kTB];
#pragma unroll
      for (int r = 0; r < kTB; ++r) d[r] = 0.f;
#pragma unroll
      for (int wi = 0; wi < 8; ++wi) {
        if (wi < words) {
#pragma unroll
          for (int bb = 0; bb < 4; ++bb) {
            const int k = wi * 4 + bb;
            if (k == 0) continue;  // byte 0 is the count
            if (k <= n) {
              const float* col = xs + ((lw[wi] >> (8 * bb)) & 0xffu);
#pragma unroll
              for (int r = 0; r < kTB; ++r) d[r] += col[r * Cp];
            }
          }
        }
      }
      if (valid) {
        float* e0 = w.emis + ((tb0 * p.B + b) * p.Lmax + st);
        const int64_t estride = p.B * p.Lmax;
#pragma unroll
        for (int r = 0; r < kTB; ++r)
          if (r < nrow) e0[r * estride] = d[r] * invC;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------ K2: lattice on emission tiles
// NS = 2: at most 72 registers, so that 28 warps share an SM (4096 sequences are then a single wave on 148 SMs)
template <int NS>
__global__ void __launch_bounds__(kLatWarps * 32, NS == 2 ? 7 : 1) lattice_tile_kernel(Problem p, TiledWs w) {
  if (*w.flag != 0) return;
  using namespace stream;
  constexpr int W = 16, TT = 8, Lpad = 16 * NS, PS = Lpad + 8, AS = Lpad + 8;
  constexpr int NPF = TT * Lpad / 32;  // emission-tile floats per lane
  constexpr int kWarpBytes = TT * PS * 4 + 2 * TT * AS * 8 + 16;
  extern __shared__ __align__(16) unsigned char smraw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * kLatWarps + warp;
  if (b >= p.B) return;
  float* pt = reinterpret_cast<float*>(smraw + (size_t)warp * kWarpBytes);
  double* abt = reinterpret_cast<double*>(pt + TT * PS);
  double* s2p = abt + 2 * TT * AS;
  const int64_t Tb64 = p.in_len[b], Lb64 = p.tgt_len[b];
  if (!seq_feasible(Tb64, Lb64, p.T, p.Lmax)) {
    if (lane == 0) p.loss[b] = INFINITY;
    return;
  }
  const int Tb = (int)Tb64, Lb = (int)Lb64, Lmax = (int)p.Lmax;
  const int NTb = (Tb + TT - 1) / TT;
  double* ck = w.ckpt + ((size_t)b * w.NT) * Lpad + (lane & (W - 1));
  int* cke = w.cke + (size_t)b * w.NT;

  // emission tile k -> registers (issued one tile ahead of its use), then -> p-tile in shared memory:
  // p_t(s) = exp(e_t(s) - rowc_t) <= 1 (floored like the fused kernel), 0 for states >= L_b and steps >= T_b
  constexpr int NPR = Lpad / 32;  // floats per lane and tile row
  const int64_t estride = p.B * (int64_t)Lmax;
  float* const e_b = w.emis + b * Lmax + lane;
  const float* const rc_b = w.rowc + b;
  bool sv[NPR];
#pragma unroll
  for (int q = 0; q < NPR; ++q) sv[q] = lane + 32 * q < Lb;
  float ev[NPF], rcv = 0.f;
  auto fetch = [&](int k) {
    const int t0 = k * TT, nv = min(TT, Tb - t0);
    rcv = lane < nv ? __ldg(rc_b + (int64_t)(t0 + lane) * p.B) : 0.f;
    const float* pe = e_b + (int64_t)t0 * estride;
#pragma unroll
    for (int r = 0; r < TT; ++r) {
#pragma unroll
      for (int q = 0; q < NPR; ++q) ev[r * NPR + q] = (r < nv && sv[q]) ? __ldg(pe + 32 * q) : 0.f;
      pe += estride;
    }
  };
  auto stage = [&](int k) {
    const int nv = min(TT, Tb - k * TT);
#pragma unroll
    for (int r = 0; r < TT; ++r) {
      const float rc = __shfl_sync(0xffffffffu, rcv, r);
#pragma unroll
      for (int q = 0; q < NPR; ++q)
        pt[r * PS + lane + 32 * q] = (r < nv && sv[q]) ? fmaxf(ex2f((ev[r * NPR + q] - rc) * kLog2e), kPMin) : 0.f;
    }
  };

  ChainScal chain;
  double cx[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) cx[j] = 0.0;
  chain.carry = (lane == 0) ? 1.0 : 0.0;
  chain.zinv = 0.0;
  chain.Ea = 0; chain.Eb = 0; chain.Ez = 0;
  // ---- phase 1: alpha, one checkpoint per tile
  fetch(0);
  for (int k = 0; k < NTb; ++k) {
    stage(k);
    if (k + 1 < NTb) fetch(k + 1);
    else if (p.grad != nullptr) fetch(NTb - 1);  // phase 2 starts with the last tile again
    __syncwarp();
    chain_phase1<NS, W, TT, PS>(cx, chain, lane, Tb, ck, cke, k, pt);
    __syncwarp();
  }
  chain_readout<NS, W>(cx, chain, lane, Lb, &p.loss[b], 1.f, nullptr);
  if (p.grad == nullptr) return;
  // ---- phase 2: tiles downwards; lanes 0-15 replay alpha from the checkpoint, lanes 16-31 run beta
  const bool isb = lane >= 16;
  for (int k = NTb - 1; k >= 0; --k) {
    stage(k);
    double ckv[NS];
    int EaK = 0;
    if (k > 0) {
#pragma unroll
      for (int j = 0; j < NS; ++j) ckv[j] = ck[(k * NS + j) * W];
      EaK = cke[k];
    } else {
#pragma unroll
      for (int j = 0; j < NS; ++j) ckv[j] = 0.0;
    }
    if (k > 0) fetch(k - 1);
    __syncwarp();
    const int Eb_all = __shfl_sync(0xffffffffu, chain.Eb, 16);
    chain_phase2<NS, W, TT, PS, AS>(cx, chain, lane, isb, Eb_all, Tb, ckv, EaK, k, pt, abt, s2p);
    __syncwarp();
    // gamma_t(s) = alpha_t(s) beta_t(s) / Z (s2 = -1/Z and the tile exponents) -> emission tile
    const double s2 = *s2p;
    const int t0 = k * TT, nv = min(TT, Tb - t0);
    float* pe = e_b + (int64_t)t0 * estride;
#pragma unroll
    for (int r = 0; r < TT; ++r) {
#pragma unroll
      for (int q = 0; q < NPR; ++q) {
        const int st = lane + 32 * q;
        if (r < nv && sv[q]) pe[32 * q] = -(float)(abt[r * AS + st] * (abt[(TT + r) * AS + st] * s2));
      }
      pe += estride;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------ K3: gradient
template <int NCI, int Lp>
__global__ void __launch_bounds__(kRowWarps * 32) bin_grad_kernel(Problem p, TiledWs w) {
  if (*w.flag != 0) return;
  extern __shared__ float smf[];  // [kRowWarps][kTB][Lp]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = blockIdx.x;
  const int64_t Tb = p.in_len[b], Lb64 = p.tgt_len[b];
  const int C = (int)p.C;
  const int64_t t0 = (int64_t)blockIdx.y * kTCh;
  const int64_t t1 = min(t0 + kTCh, p.T);
  const bool ok = seq_feasible(Tb, Lb64, p.T, p.Lmax) && (p.loss[b] < INFINITY);
  const int64_t tlive = ok ? min(t1, Tb) : t0;  // rows [tlive, t1) are zeros
  for (int64_t t = max(t0, tlive) + warp; t < t1; t += kRowWarps) {
    float* g = p.grad + (t * p.B + b) * C;
    for (int c = lane; c < C; c += 32) g[c] = 0.f;
  }
  if (!ok || t0 >= Tb) return;
  const int Lb = (int)Lb64, LW = w.LW, Lmax = (int)p.Lmax;
  const float wC = p.w_scalar * (p.seq_w ? p.seq_w[b] : 1.f) / (float)C;
  float* gs = smf + (size_t)warp * kTB * Lp;
  const int64_t rstride = p.B * p.C;
  const uint32_t* cm = w.cmask + (size_t)b * C * LW;
  // requests run one batch ahead: the next batch's logits and gamma rows are in flight while the current one is
  // walked and stored
  float xn[NCI][kTB], gn[kTB];
  auto request = [&](int64_t tb0) {
    const int nrow = (int)min((int64_t)kTB, tlive - tb0);
    const float* gam = w.emis + (tb0 * p.B + b) * Lmax + lane;
#pragma unroll
    for (int r = 0; r < kTB; ++r) gn[r] = (r < nrow && lane < Lb) ? __ldg(gam + (int64_t)r * p.B * Lmax) : 0.f;
    const float* x0 = p.logits + (tb0 * p.B + b) * p.C;
#pragma unroll
    for (int i = 0; i < NCI; ++i) {
      const int c = lane + 32 * i;
#pragma unroll
      for (int r = 0; r < kTB; ++r) xn[i][r] = (c < C && r < nrow) ? __ldg(x0 + r * rstride + c) : 0.f;
    }
  };
  const int64_t tfirst = t0 + (int64_t)warp * kTB;
  if (tfirst < tlive) request(tfirst);
  for (int64_t tb0 = tfirst; tb0 < tlive; tb0 += kRowWarps * kTB) {
    const int nrow = (int)min((int64_t)kTB, tlive - tb0);
#pragma unroll
    for (int r = 0; r < kTB; ++r) gs[r * Lp + lane] = gn[r];
    if (Lb > 32) {  // states beyond the first 32 are not prefetched
      const float* gam = w.emis + (tb0 * p.B + b) * Lmax;
      for (int s = lane + 32; s < Lb; s += 32) {
#pragma unroll
        for (int r = 0; r < kTB; ++r) gs[r * Lp + s] = r < nrow ? __ldg(gam + (int64_t)r * p.B * Lmax + s) : 0.f;
      }
    }
    float acc[NCI][kTB];
#pragma unroll
    for (int i = 0; i < NCI; ++i) {
#pragma unroll
      for (int r = 0; r < kTB; ++r) acc[i][r] = __fdividef(1.f, 1.f + __expf(-xn[i][r]));
    }
    if (tb0 + kRowWarps * kTB < tlive) request(tb0 + kRowWarps * kTB);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NCI; ++i) {
      const int c = lane + 32 * i;
      for (int wd = 0; wd < LW; ++wd) {
        uint32_t m = c < C ? __ldg(cm + (size_t)c * LW + wd) : 0u;
        const float* gw = gs + wd * 32;
        while (m) {
          const float* col = gw + (__ffs(m) - 1);
          m &= m - 1u;
#pragma unroll
          for (int r = 0; r < kTB; ++r) acc[i][r] -= col[r * Lp];
        }
      }
    }
    float* g0 = p.grad + (tb0 * p.B + b) * p.C;
#pragma unroll
    for (int i = 0; i < NCI; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
#pragma unroll
        for (int r = 0; r < kTB; ++r)
          if (r < nrow) g0[r * rstride + c] = wC * acc[i][r];
      }
    }
    __syncwarp();
  }
}

template <int NS>
int launch_lattice(const Problem& p, const TiledWs& w, cudaStream_t stream) {
  constexpr int Lpad = 16 * NS, PS = Lpad + 8, AS = Lpad + 8;
  constexpr int kWarpBytes = 8 * PS * 4 + 2 * 8 * AS * 8 + 16;
  const size_t smem = (size_t)kLatWarps * kWarpBytes;
  auto kern = lattice_tile_kernel<NS>;
  if (smem > 48 * 1024) NBCTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)((p.B + kLatWarps - 1) / kLatWarps), kLatWarps * 32, smem, stream>>>(p, w);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

template <int NCI, int Lp>
int launch_grad_inst(const Problem& p, const TiledWs& w, dim3 grid, cudaStream_t stream) {
  const size_t smem = sizeof(float) * kRowWarps * kTB * Lp;
  auto kern = bin_grad_kernel<NCI, Lp>;
  if (smem > 48 * 1024) NBCTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kRowWarps * 32, smem, stream>>>(p, w);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}
template <int NCI>
int launch_grad(const Problem& p, const TiledWs& w, dim3 grid, cudaStream_t stream) {
  switch (w.Lpad) {
    case 32: return launch_grad_inst<NCI, 32>(p, w, grid, stream);
    case 64: return launch_grad_inst<NCI, 64>(p, w, grid, stream);
    case 128: return launch_grad_inst<NCI, 128>(p, w, grid, stream);
    default: return launch_grad_inst<NCI, 256>(p, w, grid, stream);
  }
}
template <int NCI>
int launch_emis(const Problem& p, const TiledWs& w, dim3 grid, cudaStream_t stream) {
  const size_t smem = sizeof(float) * kRowWarps * kTB * (32 * NCI + 8);
  auto kern = bin_emis_kernel<NCI>;
  if (smem > 48 * 1024) NBCTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kRowWarps * 32, smem, stream>>>(p, w);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

}  // namespace

bool tiled_bin_supported(int64_t T, int64_t B, int64_t C, int64_t Lmax) {
  // class indices are bytes; the chain instances cover Lmax <= 256; grid.y = time chunks
  return C >= 1 && C <= 256 && Lmax <= 256 && B <= 0x7fffffff && (T + kTCh - 1) / kTCh <= 65535;
}

size_t tiled_bin_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax) { return layout(T, B, C, Lmax).total; }

// Launches the tiled path.  *flag_out = the device flag the caller gates the generic kernels on (they run only when
// the pre-pass found targets outside this path's domain).
int tiled_bin_launch(const Problem& p, void* ws, size_t ws_bytes, cudaStream_t stream, const int** flag_out) {
  const Layout l = layout(p.T, p.B, p.C, p.Lmax);
  if (ws == nullptr || ws_bytes < l.total) {
    set_error("workspace too small: need %zu bytes, got %zu", l.total, ws_bytes);
    return NBCTC_ERR_WORKSPACE;
  }
  char* c = static_cast<char*>(ws);
  TiledWs w;
  w.flag = reinterpret_cast<int*>(c);
  w.lists = reinterpret_cast<uint32_t*>(c + l.o_lists);
  w.cmask = reinterpret_cast<uint32_t*>(c + l.o_cmask);
  w.ckpt = reinterpret_cast<double*>(c + l.o_ckpt);
  w.cke = reinterpret_cast<int*>(c + l.o_cke);
  w.rowc = reinterpret_cast<float*>(c + l.o_rowc);
  w.emis = reinterpret_cast<float*>(c + l.o_emis);
  w.LW = l.LW; w.NT = l.NT; w.Lpad = l.Lpad;
  *flag_out = w.flag;
  NBCTC_CUDA_CHECK(cudaMemsetAsync(w.flag, 0, 256, stream));
  NBCTC_CUDA_CHECK(cudaMemsetAsync(w.cmask, 0, sizeof(uint32_t) * (size_t)p.B * p.C * l.LW, stream));
  const int64_t rows = p.B * p.Lmax;
  bin_prepass_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(p, w);
  NBCTC_LAUNCH_CHECK();

  const dim3 grid((unsigned)p.B, (unsigned)((p.T + kTCh - 1) / kTCh));
  int rc;
  const int nci = (int)((p.C + 31) / 32);
  switch (nci) {
    case 1: rc = launch_emis<1>(p, w, grid, stream); break;
    case 2: rc = launch_emis<2>(p, w, grid, stream); break;
    case 3: rc = launch_emis<3>(p, w, grid, stream); break;
    case 4: rc = launch_emis<4>(p, w, grid, stream); break;
    case 5: rc = launch_emis<5>(p, w, grid, stream); break;
    case 6: rc = launch_emis<6>(p, w, grid, stream); break;
    case 7: rc = launch_emis<7>(p, w, grid, stream); break;
    default: rc = launch_emis<8>(p, w, grid, stream); break;
  }
  if (rc != NBCTC_OK) return rc;
  switch (l.Lpad) {
    case 32: rc = launch_lattice<2>(p, w, stream); break;
    case 64: rc = launch_lattice<4>(p, w, stream); break;
    case 128: rc = launch_lattice<8>(p, w, stream); break;
    default: rc = launch_lattice<16>(p, w, stream); break;
  }
  if (rc != NBCTC_OK || p.grad == nullptr) return rc;

  switch (nci) {
    case 1: return launch_grad<1>(p, w, grid, stream);
    case 2: return launch_grad<2>(p, w, grid, stream);
    case 3: return launch_grad<3>(p, w, grid, stream);
    case 4: return launch_grad<4>(p, w, grid, stream);
    case 5: return launch_grad<5>(p, w, grid, stream);
    case 6: return launch_grad<6>(p, w, grid, stream);
    case 7: return launch_grad<7>(p, w, grid, stream);
    default: return launch_grad<8>(p, w, grid, stream);
  }
}

}  // namespace nbctc
