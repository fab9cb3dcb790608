// Sequence-per-warp no-blank CTC forward+backward kernel for sm_100a (single-label variant; host side in
// nbctc_seqwarp.cu).
//
// ONE warp owns ONE sequence from the first row to the last gradient store: no CTA barrier, no shared-memory hand-off
// between roles, nothing to poll.  A CTA is one warp, so everything derived from blockIdx is provably warp-uniform.  A
// (t,b) row of the (T,B,C) logits is read with EPL coalesced 4-byte loads per lane (lane l holds classes l, l+32, ...),
// a tile of 4 time steps at a time, after one prefetch.global.L1 per tile has requested it a tile ahead; the L lattice
// states sit NS per lane in float64.  All B sequences of a batch are in flight at once (28 warps per SM at NS = 1), so
// the latency of one sequence's dependent chain is hidden by the other warps of the SM instead of by warp
// specialisation, and the whole GPU streams through the logits time step by time step.
//
//   phase 1 (t upwards)    row log-partition (NoBlankCTC.py:136): max by one warp REDUX, the sums of the 4 rows of a
//                          tile in ONE shared butterfly; emissions p_t(s) = softmax(x_t)[label_s] (NoBlankCTC.py:96-102)
//                          from a per-lane gather load; alpha step (NoBlankCTC.py:71-87) = one 64-bit shuffle + DFMA +
//                          DMUL.  Kept for phase 2, in one workspace record per sequence: the row log-partitions
//                          (4 bytes per row), one alpha checkpoint per tile (active lanes only) and the lane scales.
//   phase 2 (t downwards)  the rows again (L2 hit for the last steps, HBM for the rest: 3 passes over N bytes instead
//                          of the lock-step kernel's 2, but at streaming speed and for any T), alpha replayed inside the
//                          tile from its checkpoint, beta backwards (the reference's own backward pass is commented out
//                          at NoBlankCTC.py:113-125; autograd does it), gamma = alpha beta / Z, and the gradient row
//                          w (softmax(x) - sum_{s: label_s = c} gamma(s)) (repeated labels
//                          accumulate): the softmax row goes out with EPL coalesced stores, then the first state of
//                          every distinct label overwrites its own class with the corrected value (deterministic: the
//                          gammas of a label are summed in ascending state order by a walk over a follower list).
//
// Numerics as in the lock-step kernel: linear domain, float64, exact power-of-two scales PER LANE
// renewed every 8 steps by a prefix-max scan (a lane is scaled to its own magnitude unless larger mass is about to
// arrive from upstream), emissions floored at 2^-120.
//
// Template parameters: NS states per lane (Lmax <= 32 NS), EPL row elements per lane (32 (EPL-1) < C <= 32 EPL).
#pragma once

#include <cuda_runtime.h>

#include "common.cuh"

namespace nbctc {

struct SwParams {
  Problem p;
  // per-sequence record in the workspace (rec_bytes apart; ONE base pointer per warp, 32-bit offsets from it):
  //   [0, 4 Tp)              row log2-partitions, Tp = 4 K
  //   [o_cke, + 128 (K/2+1)) lane scales of the alpha state after the rescale in front of tile 2r
  //   [o_ckx, + 256 NS K)    alpha before the first step of tile k (k >= 1)
  char* rec;
  int64_t rec_bytes;
  int o_cke, o_ckx;
  const int* order;     // null: sequence = ticket; else the longest-first order of the prep kernel
  int* ticket;          // null: one sequence per warp (B <= warps of the grid); else the work queue
  const float* row_lse_in;  // (T,B) or null: row log-partitions supplied by the producer of the logits
  float* row_lse_out;       // (T,B) or null: row log-partitions handed to the caller
  int* floor_flag;          // [B]: set when an emission of a live state sits on the float32 floor: the sequence is redone
                            // in the log domain by the repair kernel (nbctc_logdom.cu), which overwrites its results
  int K;                // tiles = ceil(T / 4)
  int Tp;
};

int launch_seqwarp(const SwParams& P, int NS, int EPL, int grid, cudaStream_t stream);
template <int EPL>
int launch_seqwarp_epl(const SwParams& P, int NS, int grid, cudaStream_t stream);
template <int EPL>
int seqwarp_occupancy_epl(int NS);  // resident CTAs (= warps) per SM of that instance
int seqwarp_ctas_per_sm(int NS, int EPL);

#ifdef __CUDACC__
namespace sw {

constexpr int kTT = 4;          // time steps per tile = checkpoint period
constexpr int kWarps = 1;       // warps per CTA: everything derived from blockIdx is then provably warp-uniform (uniform registers)
constexpr float kL2E = 1.4426950408889634f;
constexpr float kPFloor = 7.52316385e-37f;  // 2^-120
constexpr int kSent = -(1 << 28);           // "no exponent": an all-zero lane
constexpr unsigned kFull = 0xffffffffu;

// Largest scale step between neighbouring lanes.  Between two rescales (8 steps) mass crosses at most ceil(8/NS)
// lanes and gains 2^DEC of scaled magnitude per crossing at worst: that product stays inside the float64 range.  It
// must not be below 120: a lane's own states are flushed only when they are negligible next to the floored mass
// that arrives from upstream.
template <int NS>
struct Dec {
  static constexpr int value = NS == 1 ? 122 : NS == 2 ? 244 : NS == 4 ? 480 : 900;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float redux_max(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
// exact 2^e; 0 below the normal range, 2^1023 above
__device__ __forceinline__ double pow2z(int e) {
  e = min(e, 1023);
  return e < -1022 ? 0.0 : __hiloint2double((1023 + e) << 20, 0);
}
// min(v, 2^996) on the bit pattern: one integer instruction where fmin() on doubles takes three; inf and NaN (an
// overflowed beta factor next to an alpha that flushed to zero) come back as 2^996 too, negative values pass
__device__ __forceinline__ double clamp_big(double v) {
  return __hiloint2double(min(__double2hiint(v), 0x7E300000), __double2loint(v));
}
// volatile: the order of the row requests relative to each other is ours (the compiler would hoist every load of a
// tile to its top and spill the rows)
__device__ __forceinline__ float ldg_f(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void pf_line_l1(const void* a) { asm volatile("prefetch.global.L1 [%0];" ::"l"(a)); }

// New lane scales after a prefix-max scan along the direction mass moves (UP: towards higher lanes, alpha; else
// towards lower lanes, beta): e_l = max_{k upstream of l}(A_k - DEC dist(k,l)), A_k = absolute exponent of lane k's
// largest state.  fac = 2^(e_upstream_neighbour - e_l) brings the neighbour's state into this lane's scale.
template <int NS, bool UP>
__device__ __forceinline__ void lane_rescale(double (&x)[NS], int& e, double& fac, int lane) {
  constexpr int DEC = Dec<NS>::value;
  int hi = __double2hiint(x[0]);
#pragma unroll
  for (int j = 1; j < NS; ++j) hi = max(hi, __double2hiint(x[j]));
  const int ef = hi >> 20;  // states are >= 0 here
  const int pos = UP ? lane : 31 - lane;
  int env = ((ef > 0 && ef < 0x7ff) ? e + ef - 1023 : kSent) + DEC * pos;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    // a lane without a source gets its own value back: the max is then a no-op
    const int sh = UP ? __shfl_up_sync(kFull, env, o) : __shfl_down_sync(kFull, env, o);
    env = max(env, sh);
  }
  env -= DEC * pos;
  const int en = env > kSent / 2 ? env : e;
  const double sc = pow2z(e - en);
#pragma unroll
  for (int j = 0; j < NS; ++j) x[j] *= sc;
  e = en;
  const int eu = UP ? __shfl_up_sync(kFull, en, 1) : __shfl_down_sync(kFull, en, 1);
  fac = pos == 0 ? 0.0 : pow2z(eu - en);
}

// alpha_t(s) = (alpha_{t-1}(s) + alpha_{t-1}(s-1)) p_t(s)  (NoBlankCTC.py:73-85); states NS*lane .. NS*lane+NS-1
template <int NS>
__device__ __forceinline__ void alpha_step(double (&x)[NS], const float (&p)[NS], double fac) {
  const double up = __shfl_up_sync(kFull, x[NS - 1], 1);
#pragma unroll
  for (int j = NS - 1; j >= 1; --j) x[j] = (x[j] + x[j - 1]) * (double)p[j];
  x[0] = fma(up, fac, x[0]) * (double)p[0];
}
// u_t(s) = beta_t(s) p_t(s), beta_t(s) = u_{t+1}(s) + u_{t+1}(s+1): returns beta_t in bt and advances u
template <int NS>
__device__ __forceinline__ void beta_step(double (&u)[NS], double (&bt)[NS], const float (&p)[NS], double fac) {
  const double dn = __shfl_down_sync(kFull, u[0], 1);
#pragma unroll
  for (int j = 0; j < NS - 1; ++j) bt[j] = u[j] + u[j + 1];
  bt[NS - 1] = fma(dn, fac, u[NS - 1]);
#pragma unroll
  for (int j = 0; j < NS; ++j) u[j] = bt[j] * (double)p[j];
}

// sums of 4 per-lane values over the warp in one shared butterfly: v[i] -> total of row i in every lane
__device__ __forceinline__ void warp_sum4(float (&v)[4], int lane) {
  const bool h16 = lane & 16, h8 = lane & 8;
  // halves swap: lanes with bit 4 clear keep rows 0,1 and send rows 2,3
  float k0 = h16 ? v[2] : v[0], k1 = h16 ? v[3] : v[1];
  float s0 = h16 ? v[0] : v[2], s1 = h16 ? v[1] : v[3];
  k0 += __shfl_xor_sync(kFull, s0, 16);
  k1 += __shfl_xor_sync(kFull, s1, 16);
  float q = h8 ? k1 : k0;
  const float s = h8 ? k0 : k1;
  q += __shfl_xor_sync(kFull, s, 8);
  q += __shfl_xor_sync(kFull, q, 4);
  q += __shfl_xor_sync(kFull, q, 2);
  q += __shfl_xor_sync(kFull, q, 1);
  // q = total of row (bit4 ? 2 : 0) + (bit3 ? 1 : 0)
  v[0] = __shfl_sync(kFull, q, 0);
  v[1] = __shfl_sync(kFull, q, 8);
  v[2] = __shfl_sync(kFull, q, 16);
  v[3] = __shfl_sync(kFull, q, 24);
}
__device__ __forceinline__ float warp_sum1(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

template <int NS, int EPL>
struct Seq {
  static constexpr int Lpad = 32 * NS;

  const SwParams& P;
  const int lane;
  float* gam;             // [Lpad + 1] shared, per warp: gamma exchange, slot Lpad = 0
  unsigned short* nxt;    // [Lpad + 1] shared, per warp: next state with the same label, Lpad = none
  int64_t strideT;        // floats between two time steps of a sequence
  int C;
  bool lastok;            // lane + 32 (EPL-1) < C
  int lastoff;            // element offset of the lane's last load (an in-row element for lanes past the row's end)

  __device__ __forceinline__ Seq(const SwParams& P_, int lane_, float* gam_, unsigned short* nxt_)
      : P(P_), lane(lane_), gam(gam_), nxt(nxt_) {
    strideT = P.p.B * P.p.C;
    C = (int)P.p.C;
    lastok = lane + 32 * (EPL - 1) < C;
    lastoff = lastok ? 32 * (EPL - 1) : 0;
  }

  // rp = the lane's first element of the row
  __device__ __forceinline__ void load_row(const float* rp, float (&x)[EPL]) const {
#pragma unroll
    for (int k = 0; k < EPL - 1; ++k) x[k] = ldg_f(rp + 32 * k);
    const float v = ldg_f(rp + lastoff);
    x[EPL - 1] = lastok ? v : -INFINITY;
  }
  __device__ __forceinline__ void store_row(float* gp, const float (&r)[EPL]) const {
#pragma unroll
    for (int k = 0; k < EPL - 1; ++k) gp[32 * k] = r[k];
    if (lastok) gp[32 * (EPL - 1)] = r[EPL - 1];
  }
  __device__ __forceinline__ void zero_rows(float* g, int b, int t0, int t1) const {
    float* gp = g + ((int64_t)t0 * P.p.B + b) * C + lane;
    for (int t = t0; t < t1; ++t, gp += strideT) {
#pragma unroll
      for (int k = 0; k < EPL - 1; ++k) gp[32 * k] = 0.f;
      if (lastok) gp[32 * (EPL - 1)] = 0.f;
    }
  }
  __device__ __forceinline__ float row_max(const float (&x)[EPL]) const {
    float m = x[0];
#pragma unroll
    for (int k = 1; k < EPL; ++k) m = fmaxf(m, x[k]);
    return redux_max(m);
  }
  __device__ __forceinline__ float row_expsum(const float (&x)[EPL], float nm) const {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < EPL; ++k) s += ex2f(fmaf(x[k], kL2E, nm));
    return s;
  }
  // -log2 partition of one row, on its own (remainder steps)
  __device__ __forceinline__ float row_nl(const float* rp, int64_t tB_b) const {
    if (P.row_lse_in) return -kL2E * P.row_lse_in[tB_b];
    float r0[EPL];
    load_row(rp, r0);
    const float nm = -kL2E * row_max(r0);
    return nm - lg2f(warp_sum1(row_expsum(r0, nm)));
  }
  __device__ __forceinline__ void emissions(const float (&xg)[NS], float nl, const bool (&act)[NS], float (&pe)[NS]) const {
#pragma unroll
    for (int j = 0; j < NS; ++j) pe[j] = act[j] ? fmaxf(ex2f(fmaf(xg[j], kL2E, nl)), kPFloor) : 0.f;
  }
  // sum of the gammas of one label in ascending state order, delivered to the label's first state
  __device__ __forceinline__ void combine(float (&g)[NS], const int (&nx1)[NS], int R) const {
    if (R > 0) {
      if (NS == 1) {
        gam[lane] = g[0];
      } else if (NS == 2) {
        *reinterpret_cast<float2*>(gam + 2 * lane) = make_float2(g[0], g[NS - 1]);
      } else {
#pragma unroll
        for (int j = 0; j < NS; j += 4)
          *reinterpret_cast<float4*>(gam + NS * lane + j) = make_float4(g[j], g[(j + 1) % NS], g[(j + 2) % NS], g[(j + 3) % NS]);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NS; ++j) g[j] += gam[nx1[j]];
      if (R > 1) {
        int cur[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) cur[j] = nx1[j];
        for (int rr = 1; rr < R; ++rr) {
#pragma unroll
          for (int j = 0; j < NS; ++j) {
            cur[j] = nxt[cur[j]];
            g[j] += gam[cur[j]];
          }
        }
      }
    }
    __syncwarp();  // orders the row stores before the corrected entries; frees the exchange buffer
  }

  __device__ void run(int b);
};

template <int NS, int EPL>
__device__ void Seq<NS, EPL>::run(int b) {
  const Problem& p = P.p;
  const int T = (int)p.T, B = (int)p.B;
  const int64_t Tb64 = p.in_len[b], Lb64 = p.tgt_len[b];
  bool feas = seq_feasible(Tb64, Lb64, p.T, p.Lmax) && Lb64 <= Lpad;
  // (warp reductions of warp-uniform values: the compiler then KNOWS they are uniform, keeps them in uniform registers
  // and drops the divergence guards around every shuffle inside the loops they bound)
  const int Tb = __reduce_max_sync(kFull, feas ? (int)Tb64 : 0), Lb = __reduce_max_sync(kFull, feas ? (int)Lb64 : 0);

  // ---- labels, repeated labels
  int lab[NS];
  bool act[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int s = lane * NS + j;
    act[j] = s < Lb;
    lab[j] = act[j] ? p.labels[(int64_t)b * p.Lmax + s] : 0;
    if (act[j] && (lab[j] < 0 || lab[j] >= C)) feas = false;
  }
  feas = __all_sync(kFull, feas);
  const float w = p.w_scalar * (p.seq_w ? p.seq_w[b] : 1.f);
  // slot 0 of the record's lane-scale array (scales are stored from the second rescale on) holds the lane's repair flag
  if (lane == 0) P.floor_flag[b] = 0;
  __syncwarp();  // ordered before the 1 any lane may store in phase 1
  if (!feas) {  // outside the parity domain: +inf, zero gradient (documented deviation)
    if (lane == 0) p.loss[b] = INFINITY;
    if (p.grad) zero_rows(p.grad, b, 0, T);
    return;
  }
  bool lead[NS];
  int nx1[NS];
  int R = 0;  // largest number of earlier states with the same label
  {
    int* labs = reinterpret_cast<int*>(gam);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NS; ++j) labs[lane * NS + j] = act[j] ? lab[j] : -1 - (lane * NS + j);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int s = lane * NS + j;
      int rank = 0, nx = Lpad;
      if (act[j]) {
        for (int s2 = 0; s2 < s; ++s2) rank += labs[s2] == lab[j];
        for (int s2 = Lb - 1; s2 > s; --s2) nx = labs[s2] == lab[j] ? s2 : nx;
      }
      lead[j] = act[j] && rank == 0;
      nx1[j] = nx;
      R = max(R, rank);
    }
    R = __reduce_max_sync(kFull, R);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NS; ++j) nxt[lane * NS + j] = (unsigned short)nx1[j];
    if (lane == 0) {
      nxt[Lpad] = (unsigned short)Lpad;
      gam[Lpad] = 0.f;
    }
    __syncwarp();
  }

  const bool want_grad = p.grad != nullptr && w != 0.f;
  const bool ck_lane = act[0];  // only the lanes that hold states take part in the checkpoint traffic
  const bool have_lse = P.row_lse_in != nullptr;
  const float* const row0 = p.logits + (int64_t)b * C + lane;  // row t = 0, this lane's first class
  int goff[NS];                                                 // label's class relative to the lane's first class
#pragma unroll
  for (int j = 0; j < NS; ++j) goff[j] = lab[j] - lane;
  char* const rec = P.rec + (int64_t)b * P.rec_bytes;
  float* const lse_ws = reinterpret_cast<float*>(rec);
  double* const ckx = reinterpret_cast<double*>(rec + P.o_ckx) + lane * NS;
  int* const cke = reinterpret_cast<int*>(rec + P.o_cke) + lane;

  // ================================================================ phase 1: alpha
  // start pattern (-1)^s: the first step turns it into alpha_0 = (p_0(0), 0, 0, ...) exactly, without a special case
  double x[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) x[j] = ((lane * NS + j) & 1) ? -1.0 : 1.0;
  int e = 0;
  double fac = lane == 0 ? 0.0 : 1.0;

  const int Kf = Tb / kTT;  // full tiles
  // L1 prefetch: lane = (row of a tile, 128-byte line of the row); ONE instruction requests a whole tile, pf_near tiles
  // ahead of its loads (measured: L2 prefetches further ahead, per line or as TMA bulk prefetch, only cost time)
  const bool pf_l1 = true;
  const int LW = min((C * 4 + 127) / 128 + 1, 8);  // lanes per row: lines 0, 128, ... and the row's last element
  const bool pf_lane = lane < kTT * LW;
  const int pf_ri = lane / LW, pf_off = min((lane - pf_ri * LW) * 128, C * 4 - 4);
  // byte offsets from the running row pointers (32 bits: B C < 2^26, host gate): phase 1 from the next tile's first row,
  // phase 2 from the tile's last row to the tile below
  const int strideB = (int)strideT * 4;
  const int pf_d1 = pf_ri * strideB + pf_off - 4 * lane, pf_d2 = pf_d1 - (2 * kTT - 1) * strideB;
  // (at one state per lane the kernel sits on its 72-register limit and the absolute form below, which the compiler
  // rebuilds from the kernel arguments, measures 8 % faster there than the running-pointer form; at two states per lane
  // it is the other way round)
  const char* const pf_base = reinterpret_cast<const char*>(p.logits + (int64_t)b * C) + pf_off;
  auto pf_tile = [&](int kk) { pf_line_l1(pf_base + (int64_t)(kk * kTT + pf_ri) * (strideT * 4)); };
  const float* rq = row0;  // first row of the tile
  for (int k = 0; k < Kf; ++k) {
    float xg[kTT][NS];
    float nl[kTT];  // -log2 partition of the tile's rows
    if (have_lse) {
#pragma unroll
      for (int i = 0; i < kTT; ++i) {
        nl[i] = -kL2E * P.row_lse_in[(int64_t)(k * kTT + i) * B + b];
#pragma unroll
        for (int j = 0; j < NS; ++j) xg[i][j] = ldg_f(rq + goff[j]);
        rq += strideT;
      }
    } else {
      float xr[kTT][EPL];
#pragma unroll
      for (int i = 0; i < kTT; ++i) {
        load_row(rq, xr[i]);
#pragma unroll
        for (int j = 0; j < NS; ++j) xg[i][j] = ldg_f(rq + goff[j]);
        rq += strideT;
      }
      if (pf_l1 && pf_lane && k + 1 < Kf) {
        if (NS == 1) pf_tile(k + 1);
        else pf_line_l1(reinterpret_cast<const char*>(rq) + pf_d1);
      }
      float nm[kTT], s[kTT];
#pragma unroll
      for (int i = 0; i < kTT; ++i) nm[i] = -kL2E * row_max(xr[i]);
#pragma unroll
      for (int i = 0; i < kTT; ++i) s[i] = row_expsum(xr[i], nm[i]);
      warp_sum4(s, lane);
#pragma unroll
      for (int i = 0; i < kTT; ++i) nl[i] = nm[i] - lg2f(s[i]);
    }
    float pe[kTT][NS];
#pragma unroll
    for (int i = 0; i < kTT; ++i) emissions(xg[i], nl[i], act, pe[i]);
    {  // an emission on the float32 floor: the repair kernel redoes the sequence (nothing is carried through the loop
       // for it: the kernel sits on its register limit; the rest of this sequence's work here is simply overwritten)
      bool low = false;
      if (NS == 1) {  // (a lane without a state holds zeros: never equal to the floor)
        low = fminf(fminf(pe[0][0], pe[1][0]), fminf(pe[2][0], pe[3][0])) == kPFloor;
      } else {
#pragma unroll
        for (int i = 0; i < kTT; ++i)
#pragma unroll
          for (int j = 0; j < NS; ++j) low |= pe[i][j] == kPFloor;
      }
      if (low) P.floor_flag[b] = 1;
    }
    if (want_grad && lane == 0) *reinterpret_cast<float4*>(lse_ws + k * kTT) = make_float4(-nl[0], -nl[1], -nl[2], -nl[3]);
    if (P.row_lse_out && lane < kTT) {
      const float v = lane == 0 ? nl[0] : lane == 1 ? nl[1] : lane == 2 ? nl[2] : nl[3];
      P.row_lse_out[(int64_t)(k * kTT + lane) * B + b] = v * -0.6931471805599453f;
    }
    if (k > 0) {
      if ((k & 1) == 0) {
        lane_rescale<NS, true>(x, e, fac, lane);
        if (want_grad && ck_lane) cke[(k >> 1) * 32] = e;
      }
      if (want_grad && ck_lane) {
#pragma unroll
        for (int j = 0; j < NS; ++j) ckx[k * Lpad + j] = x[j];
      }
    }
#pragma unroll
    for (int i = 0; i < kTT; ++i) alpha_step<NS>(x, pe[i], fac);
  }
  // remainder steps (Tb % 4), one row at a time
  const int nrem = Tb - Kf * kTT;
  if (nrem > 0) {
    const int k = Kf;
    if (k > 0) {
      if ((k & 1) == 0) {
        lane_rescale<NS, true>(x, e, fac, lane);
        if (want_grad && ck_lane) cke[(k >> 1) * 32] = e;
      }
      if (want_grad && ck_lane) {
#pragma unroll
        for (int j = 0; j < NS; ++j) ckx[k * Lpad + j] = x[j];
      }
    }
    for (int i = 0; i < nrem; ++i) {
      const int t = k * kTT + i;
      const float* rp = row0 + (int64_t)t * strideT;
      const float nl = row_nl(rp, (int64_t)t * B + b);
      if (want_grad && lane == 0) lse_ws[t] = -nl;
      if (P.row_lse_out && lane == 0) P.row_lse_out[(int64_t)t * B + b] = nl * -0.6931471805599453f;
      float g1[NS], pr[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) g1[j] = ldg_f(rp + goff[j]);
      emissions(g1, nl, act, pr);
      {
        bool low = false;
#pragma unroll
        for (int j = 0; j < NS; ++j) low |= pr[j] == kPFloor;
        if (low) P.floor_flag[b] = 1;
      }
      alpha_step<NS>(x, pr, fac);
    }
  }
  // ---- read-out (NoBlankCTC.py:58-68,:139): Z = alpha_{T_b-1}(L_b-1)
  double zinv;
  int Ez;
  {
    const int sl = Lb - 1, rj = sl % NS;
    double mine = x[0];
#pragma unroll
    for (int j = 1; j < NS; ++j) mine = (rj >= j) ? x[j] : mine;
    double zhat = __shfl_sync(kFull, mine, sl / NS);
    Ez = __shfl_sync(kFull, e, sl / NS);
    const int ezf = __double2hiint(zhat) >> 20;
    const bool ok = zhat > 0.0 && ezf > 0 && ezf < 0x7ff;
    if (ok) {  // normalise to [1, 2)
      zhat *= pow2z(1023 - ezf);
      Ez += ezf - 1023;
    }
    if (lane == 0) p.loss[b] = ok ? (float)(-(log(zhat) + (double)Ez * 0.6931471805599453)) : INFINITY;
    zinv = ok ? (double)w / zhat : 0.0;  // sequence weight folded into gamma
    if (!ok) {
      if (p.grad) zero_rows(p.grad, b, 0, T);
      return;
    }
  }
  if (!p.grad) return;
  if (!want_grad) {  // zero weight
    zero_rows(p.grad, b, 0, T);
    return;
  }
  __syncwarp();  // lane 0's log-partition stores are read back by every lane below

  // ================================================================ phase 2: alpha replay, beta, gradient
  const float lw = lg2f(w);
  double u[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int s = lane * NS + j;
    u[j] = s < Lb ? (((Lb - 1 - s) & 1) ? -1.0 : 1.0) : 0.0;
  }
  int eb = 0;
  double facb = lane == 31 ? 0.0 : 1.0;
  const int64_t gdelta = p.grad - p.logits;  // same layout: gradient element = logits element + gdelta

  // one backward step at time t: beta, gamma, the gradient row.  a = alpha_t scaled by ga; xrow/xgv = the row's logits.
  auto grad_step = [&](const double (&a)[NS], const float (&pe)[NS], float nl, double gb, const float (&xrow)[EPL],
                       float* grow) {
    double bt[NS];
    beta_step<NS>(u, bt, pe, facb);
    float g[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) g[j] = (float)(a[j] * clamp_big(bt[j] * gb));
    const float nlw = nl + lw;  // the softmax row, scaled by the sequence weight
    float r[EPL];
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) r[kk] = ex2f(fmaf(xrow[kk], kL2E, nlw));
    store_row(grow, r);
    combine(g, nx1, R);
#pragma unroll
    for (int j = 0; j < NS; ++j)
      if (lead[j]) grow[goff[j]] = fmaf(w, pe[j], -g[j]);  // w softmax(x)[label] = w p (the 2^-120 floor is 0 in float32 terms)
  };
  // alpha state and scale in front of tile k
  auto load_ck = [&](int k, double (&xa)[NS], int& ea) {
    ea = 0;
    if (k == 0) {
#pragma unroll
      for (int j = 0; j < NS; ++j) xa[j] = ((lane * NS + j) & 1) ? -1.0 : 1.0;
    } else {
#pragma unroll
      for (int j = 0; j < NS; ++j) xa[j] = ck_lane ? ckx[k * Lpad + j] : 0.0;
      if (k >= 2 && ck_lane) ea = cke[(k >> 1) * 32];
    }
  };
  // gamma_t(s) = alpha beta / Z: the three power-of-two scales split over two exact factors.  ga = 2^Ha <= 1 multiplies
  // the replayed alpha (after the recursion: folded into the recursion it would enter the neighbour factor, which for
  // lanes whose Ha differ by ~1000 leaves the float64 range), gb = 2^(H-Ha) / Z multiplies beta.
  auto gscales = [&](int ea, double& ga, double& gb, double& faca) {
    const int H = ea + eb - Ez;
    const int Ha = max(min(H, 0), -1000);
    ga = pow2z(Ha);
    gb = pow2z(H - Ha) * zinv;
    const int eu = __shfl_up_sync(kFull, ea, 1);
    faca = lane == 0 ? 0.0 : pow2z(eu - ea);
  };

  if (nrem > 0) {  // the sequence's last steps, one at a time
    const int k = Kf;
    double xa0[NS], ga, gb, faca;
    int ea;
    load_ck(k, xa0, ea);
    gscales(ea, ga, gb, faca);
    for (int i = nrem - 1; i >= 0; --i) {
      double xa[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) xa[j] = xa0[j];
      float pr[NS], g1[NS], nl = 0.f;
      const float* rp = row0 + (int64_t)(k * kTT) * strideT;
      for (int ii = 0; ii <= i; ++ii, rp += strideT) {  // replay up to step i
        nl = -lse_ws[k * kTT + ii];
#pragma unroll
        for (int j = 0; j < NS; ++j) g1[j] = ldg_f(rp + goff[j]);
        emissions(g1, nl, act, pr);
        alpha_step<NS>(xa, pr, faca);
      }
      rp -= strideT;
      float r0[EPL];
      load_row(rp, r0);
#pragma unroll
      for (int j = 0; j < NS; ++j) xa[j] *= ga;
      grad_step(xa, pr, nl, gb, r0, const_cast<float*>(rp) + gdelta);
    }
    lane_rescale<NS, false>(u, eb, facb, lane);
  }

  if (Kf > 0) {
    // rows of the first tiles are still in L2 (phase 1 has just read them); request the ones further down
    const float* rl = row0 + (int64_t)(Kf * kTT - 1) * strideT;  // next row to load, walking down
    int since = 0;                                              // tiles since the last beta rescale
    // the tile's checkpoint, lane scale and log-partitions are loaded one tile ahead of their use
    double xa_n[NS];
    int ea_n;
    float4 l4_n;
    load_ck(Kf - 1, xa_n, ea_n);
    l4_n = *reinterpret_cast<const float4*>(lse_ws + (Kf - 1) * kTT);
    for (int k = Kf - 1; k >= 0; --k) {
      if (pf_l1 && pf_lane && k >= 1) {
        if (NS == 1) pf_tile(k - 1);
        else pf_line_l1(reinterpret_cast<const char*>(rl) + pf_d2);
      }
      float xg[kTT][NS];
      {
        const float* rg = rl;
#pragma unroll
        for (int i = kTT - 1; i >= 0; --i) {
#pragma unroll
          for (int j = 0; j < NS; ++j) xg[i][j] = ldg_f(rg + goff[j]);
          rg -= strideT;
        }
      }
      float xr[2][EPL];  // the row being worked on and the next one down
      load_row(rl, xr[(kTT - 1) & 1]);
      double xa[NS], ga, gb, faca;
#pragma unroll
      for (int j = 0; j < NS; ++j) xa[j] = xa_n[j];
      const int ea = ea_n;
      const float nl[kTT] = {-l4_n.x, -l4_n.y, -l4_n.z, -l4_n.w};
      if (k > 0) {
        load_ck(k - 1, xa_n, ea_n);
        l4_n = *reinterpret_cast<const float4*>(lse_ws + (k - 1) * kTT);
      }
      gscales(ea, ga, gb, faca);
      float pe[kTT][NS];
      double a[kTT][NS];
#pragma unroll
      for (int i = 0; i < kTT; ++i) emissions(xg[i], nl[i], act, pe[i]);
#pragma unroll
      for (int i = 0; i < kTT; ++i) {
        alpha_step<NS>(xa, pe[i], faca);
#pragma unroll
        for (int j = 0; j < NS; ++j) a[i][j] = xa[j] * ga;
      }
      float* gr = const_cast<float*>(rl) + gdelta;  // gradient row of the tile's last step
#pragma unroll
      for (int i = kTT - 1; i >= 0; --i) {
        rl -= strideT;
        if (i > 0) load_row(rl, xr[(i - 1) & 1]);
        grad_step(a[i], pe[i], nl[i], gb, xr[i & 1], gr);
        gr -= strideT;
      }
      if (++since == 2) {
        since = 0;
        lane_rescale<NS, false>(u, eb, facb, lane);
      }
    }
  }
  if (Tb < T) zero_rows(p.grad, b, Tb, T);
}

template <int NS, int EPL>
__global__ void __launch_bounds__(kWarps * 32, NS == 1 ? 28 : NS == 2 ? 20 : 12) seqwarp_kernel(const __grid_constant__ SwParams P) {
  constexpr int Lpad = 32 * NS;
  __shared__ float s_gam[kWarps][Lpad + 4];
  __shared__ unsigned short s_nxt[kWarps][Lpad + 4];
  const int lane = kWarps == 1 ? threadIdx.x : threadIdx.x & 31, wib = kWarps == 1 ? 0 : threadIdx.x >> 5;
  // the repair kernel behind this one is launched with programmatic stream serialization: its CTAs may take their places
  // as ours retire (they wait for this whole grid before they read a flag)
  asm volatile("griddepcontrol.launch_dependents;");
  Seq<NS, EPL> sq(P, lane, s_gam[wib], s_nxt[wib]);
  const int B = (int)P.p.B;
  int i = blockIdx.x * kWarps + wib;
  for (;;) {
    if (P.ticket) {
      if (lane == 0) i = atomicAdd(P.ticket, 1);
      i = __shfl_sync(kFull, i, 0);
    }
    if (i >= B) break;
    sq.run(P.order ? P.order[i] : i);
    i += gridDim.x * kWarps;
  }
}

template <int NS, int EPL>
int launch_one(const SwParams& P, int grid, cudaStream_t stream) {
  seqwarp_kernel<NS, EPL><<<grid, kWarps * 32, 0, stream>>>(P);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

}  // namespace sw

template <int EPL>
int seqwarp_occupancy_epl(int NS) {
  int n = 0;
  cudaError_t e = NS == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sw::seqwarp_kernel<1, EPL>, sw::kWarps * 32, 0)
                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sw::seqwarp_kernel<2, EPL>, sw::kWarps * 32, 0);
  return e == cudaSuccess ? n : 0;
}

template <int EPL>
int launch_seqwarp_epl(const SwParams& P, int NS, int grid, cudaStream_t stream) {
  switch (NS) {
    case 1: return sw::launch_one<1, EPL>(P, grid, stream);
    case 2: return sw::launch_one<2, EPL>(P, grid, stream);
    default: set_error("seqwarp: NS=%d not built", NS); return NBCTC_ERR_UNSUPPORTED;
  }
}
#endif  // __CUDACC__

}  // namespace nbctc
