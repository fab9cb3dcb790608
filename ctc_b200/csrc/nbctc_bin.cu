// Tiled path of the multi-label variant (NoBlankBinaryCTC.py): three kernels, time-batched row kernels.
//
// The multi-hot targets y[b] (L_b x C) do not change over time, so every product with them is organised as
// "decode an index once, use it for kTB = 8 time steps":
//
//   K1 bin_emis      : CTA = one sequence x 256 time steps.  First the class lists of the sequence's states (one warp
//                      per target row: <= 31 byte indices) and, from the first chunk's CTA, the class -> states bit
//                      masks; a row that is not exact {0,1} or holds more than 31 classes raises a flag -- the
//                      kernels below then return at once and the generic kernels (gated on the same flag) redo the
//                      call.  Then warp = batches of 8 rows staged in shared memory;
//                      emissions p_t(s) = exp(e_t(s)), e_t(s) = (1/C) (sum_{c in S_s} x_t(c) - sum_c softplus(x_t(c)))
//                      with lane = state (NoBlankBinaryCTC.py:109-112: e = -BCELoss(sigmoid(x_t), y_s)).
//   K2 lattice_tile  : warp = sequence; the float64 linear-domain chain of the fused kernel (stream_kernel.cuh:
//                      16 lanes x NS states per direction, exact power-of-two rescaling, one alpha checkpoint per
//                      tile of 8 steps, alpha replay next to beta in phase 2) on the emission tiles;
//                      gamma overwrites the emissions (NoBlankBinaryCTC.py:72-95 transition, read-out :58-68).
//   K12 bin_seq      : Lmax <= 32 (BASELINE configs[2]): K1 and K2 as ONE sequence-per-warp kernel (the scheme of
//                      seqwarp_kernel.cuh: warp = sequence, tiles of 4 steps, float64 chain on all 32 lanes with
//                      per-lane scales, alpha replay from a checkpoint per tile) -- see its comment below.
//   K3 bin_grad      : CTA = one sequence x 256 time steps, 16 warps, warp = batches of 4 rows staged by TMA bulk
//                      copies (double buffered), lane = class: grad = w/C * (sigmoid(x) - sum_{s in M_c} gamma_t(s)),
//                      the states of a class decoded once per CTA into a round table (ascending state order:
//                      deterministic).
//
// HBM traffic: logits read twice, gradient written once, emission/gamma tile (T,B,Lmax) fp32 written twice and read
// twice, checkpoints (T/8,B,Lpad) f64 -- about 1.8x the algorithmic bytes at C = 157, Lmax = 32.
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "seqwarp_kernel.cuh"  // sw:: chain steps, lane scales, reductions
#include "stream_kernel.cuh"

namespace nbctc {

namespace {

constexpr int kTB = 8;        // rows (time steps) per warp batch
constexpr int kRowWarps = 8;  // warps per CTA in K1 / K3
constexpr int kTCh = 256;     // time steps per CTA in K1 / K3
constexpr int kLatWarps = 4;  // sequences per CTA in K2

struct TiledWs {
  int* flag;            // != 0: the targets are outside this path's domain -> generic kernels
  uint32_t* cmask;      // [B][C][LW]    states that contain the class
  double* ckpt;         // [B][NT][Lpad] alpha checkpoints
  int* cke;             // [B][NT][16]   their lane scales
  float* emis;          // (T,B,Lmax)    emissions p_t(s), overwritten by gamma
  char* rec;            // K12 (Lmax <= 32): per sequence [K/2+1][32] int lane scales | [K][32] float64 alpha checkpoints
  int64_t rec_bytes;
  int o_rec_ckx;
  int LW, NT, Lpad;
  int rs32, es32;       // K12: B C and B Lmax, the row strides of logits and emission tile in floats (host gate: < 2^26)
};

struct Layout {
  size_t o_cmask, o_ckpt, o_cke, o_emis, o_rec, rec_bytes, total;
  int o_rec_ckx;
  bool seq;  // Lmax <= 32: emissions and lattice in ONE sequence-per-warp kernel (K12) instead of K1 + K2
  int LW, NT, Lpad;
};

Layout layout(int64_t T, int64_t B, int64_t C, int64_t Lmax) {
  Layout l{};
  l.Lpad = Lmax <= 32 ? 32 : Lmax <= 64 ? 64 : Lmax <= 128 ? 128 : 256;
  l.LW = (int)((Lmax + 31) / 32);
  l.NT = (int)((T + 7) / 8);
  size_t off = 256;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  l.o_cmask = take(sizeof(uint32_t) * (size_t)B * C * l.LW);
  l.o_ckpt = take(sizeof(double) * (size_t)B * l.NT * l.Lpad);
  l.o_cke = take(sizeof(int) * (size_t)B * l.NT * 16);
  l.o_emis = take(sizeof(float) * (size_t)T * B * Lmax);
  l.seq = Lmax <= 32 && T < (1 << 24) && B * C < ((int64_t)1 << 26) && B * Lmax < ((int64_t)1 << 26);  // (32-bit row strides in the kernel)
  if (l.seq) {
    const size_t K = (size_t)((T + 3) / 4);
    l.o_rec_ckx = (int)align_up(sizeof(int) * (K / 2 + 1) * 32, 256);
    l.rec_bytes = align_up(l.o_rec_ckx + sizeof(double) * K * 32, 256);
    l.o_rec = take(l.rec_bytes * (size_t)B);
  }
  l.total = off;
  return l;
}

// 1 / (1 + exp(-v)) with the approximate ex2 and rcp units (2 ulp): 4 instructions
__device__ __forceinline__ float sigmoid_fast(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + stream::ex2f(-v * stream::kLog2e)));
  return r;
}
__device__ __forceinline__ float lg2f_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------ K1: emissions
template <int NCI>
__global__ void __launch_bounds__(kRowWarps * 32, NCI <= 5 ? 4 : 2) bin_emis_kernel(Problem p, TiledWs w) {
  constexpr int Cp = 32 * NCI + 8;  // row stride: rows start 8 banks apart
  extern __shared__ float smf[];  // [kRowWarps][kTB][Cp] logits rows | [Lmax][8] class lists
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
  // Class lists of the sequence's states (one warp per target row): byte 0 = class count (<= 31), bytes 1.. = class
  // indices.  A row that is not exact {0,1} or holds more than 31 classes raises the flag: the later kernels then
  // return at once and the gated generic kernels redo the whole call.  The CTA of the sequence's first time chunk
  // also publishes the class -> states bit masks for the gradient kernel.
  uint32_t* sl = reinterpret_cast<uint32_t*>(smf + (size_t)kRowWarps * kTB * Cp);
  {
    const bool publish = blockIdx.y == 0;
    for (int s = warp; s < Lb; s += kRowWarps) {
      const float* y = p.targets + ((size_t)b * p.Lmax + s) * C;
      float yv[NCI];
#pragma unroll
      for (int i = 0; i < NCI; ++i) yv[i] = (i + 1 < NCI || lane + 32 * i < C) ? __ldg(y + lane + 32 * i) : 0.f;
      uint32_t* rec = sl + s * 8;
      if (lane < 8) rec[lane] = 0u;
      __syncwarp();
      unsigned char* bytes = reinterpret_cast<unsigned char*>(rec);
      int count = 0;
      bool bad = false;
#pragma unroll
      for (int i = 0; i < NCI; ++i) {
        const int c = lane + 32 * i;
        const bool one = yv[i] == 1.f;
        bad |= !(one || yv[i] == 0.f);
        const unsigned m = __ballot_sync(0xffffffffu, one);
        const int pos = count + __popc(m & ((1u << lane) - 1u));
        if (one && pos < 31) bytes[1 + pos] = (unsigned char)c;
        if (one && publish) atomicOr(&w.cmask[((size_t)b * C + c) * w.LW + (s >> 5)], 1u << (s & 31));
        count += __popc(m);
      }
      bad = __any_sync(0xffffffffu, bad) || count > 31;
      if (lane == 0) {
        bytes[0] = (unsigned char)min(count, 31);
        if (bad) atomicOr(w.flag, 1);
      }
    }
    __syncthreads();
  }
  // the whole batch of a warp is in flight before its first use: one memory round trip per batch.  (Requesting the
  // next batch ahead of the walk keeps v live across it: 124 registers, two CTAs per SM, 0.34 instead of 0.26 ms.)
  float v[NCI][kTB];
  auto request = [&](int64_t tb0) {
    const int nrow = (int)min((int64_t)kTB, tend - tb0);
    const float* xr = p.logits + (tb0 * p.B + b) * p.C + lane;
#pragma unroll
    for (int r = 0; r < kTB; ++r) {
#pragma unroll
      for (int i = 0; i < NCI; ++i) v[i][r] = ((i + 1 < NCI || lane + 32 * i < C) && r < nrow) ? __ldg(xr + 32 * i) : 0.f;
      xr += rstride;
    }
  };
  for (int64_t tb0 = t0 + (int64_t)warp * kTB; tb0 < tend; tb0 += kRowWarps * kTB) {
    const int nrow = (int)min((int64_t)kTB, tend - tb0);
    request(tb0);
    // sum_c softplus(x_c) = sum_c max(x_c, 0) + log prod_c (1 + exp(-|x_c|)): every factor is in (1, 2], so the
    // product of a lane's NCI <= 8 factors cannot overflow and one lg2 per lane and row replaces one per element
    float sp[kTB];
    {
      float pos[kTB], prod[kTB];
#pragma unroll
      for (int r = 0; r < kTB; ++r) { pos[r] = 0.f; prod[r] = 1.f; }
#pragma unroll
      for (int i = 0; i < NCI; ++i) {
        const int c = lane + 32 * i;
        const bool in = i + 1 < NCI || c < C;
#pragma unroll
        for (int r = 0; r < kTB; ++r) {
          const float x = v[i][r];
          xs[r * Cp + c] = x;
          const float f = 1.f + stream::ex2f(-fabsf(x) * stream::kLog2e);
          pos[r] += in ? fmaxf(x, 0.f) : 0.f;
          prod[r] *= in ? f : 1.f;
        }
      }
#pragma unroll
      for (int r = 0; r < kTB; ++r) sp[r] = fmaf(lg2f_fast(prod[r]), 0.6931471805599453f, pos[r]);
    }
#pragma unroll
    for (int r = 0; r < kTB; ++r) sp[r] = warp_sum(sp[r]);
    __syncwarp();
    for (int s0 = 0; s0 < Lb; s0 += 32) {
      const int st = s0 + lane;
      const bool valid = st < Lb;
      uint32_t lw[8];
      {
        const uint4* src = reinterpret_cast<const uint4*>(sl + (size_t)min(st, Lb - 1) * 8);
        const uint4 a = src[0], c4 = src[1];
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
      // p_t(s) = exp((sum_{c in S_s} x_c - sum_c softplus(x_c)) / C) <= 1, floored like the fused kernel's emissions;
      // zeros for the states [L_b, Lmax) of the block (the lattice kernel loads whole pairs)
      if (st < p.Lmax) {
        float* e0 = w.emis + ((tb0 * p.B + b) * p.Lmax + st);
        const int64_t estride = p.B * p.Lmax;
        const float k2 = invC * stream::kLog2e;
#pragma unroll
        for (int r = 0; r < kTB; ++r)
          if (r < nrow) e0[r * estride] = valid ? fmaxf(stream::ex2f((d[r] - sp[r]) * k2), stream::kPMin) : 0.f;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------ K12: emissions + lattice, one warp per sequence
// Lmax <= 32 (BASELINE configs[2]): the sequence-per-warp scheme of seqwarp_kernel.cuh applied to the multi-label
// variant.  One warp owns one sequence: class lists once; then, tile of 4 time steps by tile, the rows' softplus sums and
// the emissions p_t(s) (lane = state; a class byte is decoded once and used for the 4 rows of the tile), the alpha
// step in float64 with per-lane scales, p_t(s) parked in the (T,B,Lmax) tile; read-out; then downwards: p back from the
// tile, alpha replayed from the tile's checkpoint, beta, gamma = alpha beta / Z over p in the tile -- what K3 consumes.
// Against K1 + K2: the emission tile is written once and read once less, the lattice runs on all 32 lanes instead of
// 16 + 16, and there is one launch and one pass over the label bookkeeping instead of two.
template <int NCI>
__global__ void __launch_bounds__(32, 28) bin_seq_kernel(Problem p, TiledWs w) {
  using namespace sw;
  static_assert(kTT == 4, "a class's values of the tile's rows are one float4");
  __shared__ float4 xs[32 * NCI];  // [class] -> the class's logits in the tile's 4 rows: one STS.128 / LDS.128 per class
  __shared__ __align__(16) uint32_t sl[32 * 8];
  const int lane = threadIdx.x;
  const int64_t b = blockIdx.x;
  const int64_t Tb64 = p.in_len[b], Lb64 = p.tgt_len[b];
  const bool feas = seq_feasible(Tb64, Lb64, p.T, p.Lmax);
  const int Tb = __reduce_max_sync(kFull, feas ? (int)Tb64 : 0), Lb = __reduce_max_sync(kFull, feas ? (int)Lb64 : 0);
  if (Tb == 0) {
    if (lane == 0) p.loss[b] = INFINITY;
    return;
  }
  const int C = (int)p.C, Lmax = (int)p.Lmax;
  // ---- class lists of the states (as K1) and the class -> states masks of K3
  {
    bool any_bad = false;
    for (int s = 0; s < Lb; ++s) {
      const float* y = p.targets + ((size_t)b * Lmax + s) * C;
      float yv[NCI];
#pragma unroll
      for (int i = 0; i < NCI; ++i) yv[i] = (i + 1 < NCI || lane + 32 * i < C) ? __ldg(y + lane + 32 * i) : 0.f;
      uint32_t* rec = sl + s * 8;
      if (lane < 8) rec[lane] = 0u;
      __syncwarp();
      unsigned char* bytes = reinterpret_cast<unsigned char*>(rec);
      int count = 0;
      bool bad = false;
#pragma unroll
      for (int i = 0; i < NCI; ++i) {
        const int c = lane + 32 * i;
        const bool one = yv[i] == 1.f;
        bad |= !(one || yv[i] == 0.f);
        const unsigned m = __ballot_sync(kFull, one);
        const int pos = count + __popc(m & ((1u << lane) - 1u));
        if (one && pos < 31) bytes[1 + pos] = (unsigned char)c;
        if (one) atomicOr(&w.cmask[((size_t)b * C + c) * w.LW], 1u << s);
        count += __popc(m);
      }
      any_bad |= __any_sync(kFull, bad) || count > 31;
      if (lane == 0) bytes[0] = (unsigned char)min(count, 31);
    }
    __syncwarp();
    if (any_bad) {  // outside this path's domain: the gated generic kernels redo the call
      if (lane == 0) atomicOr(w.flag, 1);
      return;
    }
  }
  const bool valid = lane < Lb;
  uint32_t lw[8];
  {
    const uint4* src = reinterpret_cast<const uint4*>(sl + (size_t)min(lane, Lb - 1) * 8);
    const uint4 a = src[0], c4 = src[1];
    lw[0] = a.x; lw[1] = a.y; lw[2] = a.z; lw[3] = a.w;
    lw[4] = c4.x; lw[5] = c4.y; lw[6] = c4.z; lw[7] = c4.w;
  }
  const int n = valid ? (int)(lw[0] & 0xffu) : 0;
  const int words = (__reduce_max_sync(kFull, n) + 4) >> 2;  // bytes 0..n
  const float k2 = stream::kLog2e / (float)C;

  // row strides as 32-bit kernel arguments: used straight from the constant bank (the 64-bit products p.B * p.C were
  // rebuilt in every tile: a fifth of the kernel's instructions went into addresses)
  const int rstride = w.rs32, estride = w.es32;
  float* const e_b = w.emis + b * Lmax + lane;
  char* const rec = w.rec + b * w.rec_bytes;
  int* const cke = reinterpret_cast<int*>(rec) + lane;
  double* const ckx = reinterpret_cast<double*>(rec + w.o_rec_ckx) + lane;
  const bool estore = lane < Lmax;
  const bool want_grad = p.grad != nullptr;

  // ================================================================ phase 1
  double x[1] = {(lane & 1) ? -1.0 : 1.0};
  int e = 0;
  double fac = lane == 0 ? 0.0 : 1.0;
  const int K = (Tb + kTT - 1) / kTT;
  const float* xr = p.logits + b * p.C + lane;  // row t = 0
  // one prefetch.global.L1 per tile requests the next tile's rows: lane = (row of the tile, 128-byte line of the row)
  const int LW = min((C * 4 + 127) / 128 + 1, 8);
  const int pf_ri = lane / LW;
  const int pf_d = pf_ri * rstride * 4 + min((lane - pf_ri * LW) * 128, C * 4 - 4) - 4 * lane;
  const bool pf_lane = lane < kTT * LW;
  float* ep1 = e_b;  // the tile's first emission row
  // one tile; F: all kTT rows are live (every tile but the sequence's last: no per-row predicates)
  auto tile1 = [&](int k, auto full) {
    constexpr bool F = decltype(full)::value;
    const int nrow = F ? kTT : Tb - k * kTT;
    float v[NCI][kTT];
#pragma unroll
    for (int r = 0; r < kTT; ++r) {
#pragma unroll
      for (int i = 0; i < NCI; ++i) v[i][r] = ((i + 1 < NCI || lane + 32 * i < C) && (F || r < nrow)) ? ldg_f(xr + 32 * i) : 0.f;
      xr += rstride;
    }
    if (pf_lane && (k + 1) * kTT + pf_ri < Tb) pf_line_l1(reinterpret_cast<const char*>(xr) + pf_d);
    // sum_c softplus(x_c) = sum_c max(x_c, 0) + log prod_c (1 + exp(-|x_c|))  (K1)
    float sp[kTT];
    {
      float pos[kTT], prod[kTT];
#pragma unroll
      for (int r = 0; r < kTT; ++r) { pos[r] = 0.f; prod[r] = 1.f; }
#pragma unroll
      for (int i = 0; i < NCI; ++i) {
        const int c = lane + 32 * i;
        const bool in = i + 1 < NCI || c < C;
        xs[c] = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
#pragma unroll
        for (int r = 0; r < kTT; ++r) {
          const float xv = v[i][r];
          const float f = 1.f + stream::ex2f(-fabsf(xv) * stream::kLog2e);
          pos[r] += in ? fmaxf(xv, 0.f) : 0.f;
          prod[r] *= in ? f : 1.f;
        }
      }
#pragma unroll
      for (int r = 0; r < kTT; ++r) sp[r] = fmaf(lg2f_fast(prod[r]), 0.6931471805599453f, pos[r]);
    }
    warp_sum4(sp, lane);
    __syncwarp();
    float d[kTT];
#pragma unroll
    for (int r = 0; r < kTT; ++r) d[r] = 0.f;
#pragma unroll
    for (int wi = 0; wi < 8; ++wi) {
      if (wi < words) {
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          const int kk = wi * 4 + bb;
          if (kk == 0) continue;  // byte 0 is the count
          if (kk <= n) {
            const float4 q = xs[(lw[wi] >> (8 * bb)) & 0xffu];
            d[0] += q.x; d[1] += q.y; d[2] += q.z; d[3] += q.w;
          }
        }
      }
    }
    float pe[kTT][1];
#pragma unroll
    for (int r = 0; r < kTT; ++r) pe[r][0] = valid ? fmaxf(stream::ex2f((d[r] - sp[r]) * k2), stream::kPMin) : 0.f;
    if (want_grad && estore) {
      float* e0 = ep1;
#pragma unroll
      for (int r = 0; r < kTT; ++r) {
        if (F || r < nrow) *e0 = pe[r][0];
        e0 += estride;
      }
    }
    ep1 += kTT * estride;
    __syncwarp();  // the staged rows are free for the next tile
    if (k > 0) {
      if ((k & 1) == 0) {
        lane_rescale<1, true>(x, e, fac, lane);
        if (want_grad && valid) cke[(k >> 1) * 32] = e;
      }
      if (want_grad && valid) ckx[k * 32] = x[0];
    }
#pragma unroll
    for (int r = 0; r < kTT; ++r)
      if (F || r < nrow) alpha_step<1>(x, pe[r], fac);
  };
  const int Kf = Tb / kTT;  // full tiles
  for (int k = 0; k < Kf; ++k) tile1(k, std::true_type{});
  if (Kf < K) tile1(Kf, std::false_type{});
  // ---- read-out (NoBlankBinaryCTC.py:58-68)
  double zinv;
  int Ez;
  {
    double zhat = __shfl_sync(kFull, x[0], Lb - 1);
    Ez = __shfl_sync(kFull, e, Lb - 1);
    const int ezf = __double2hiint(zhat) >> 20;
    const bool ok = zhat > 0.0 && ezf > 0 && ezf < 0x7ff;
    if (ok) {
      zhat *= pow2z(1023 - ezf);
      Ez += ezf - 1023;
    }
    if (lane == 0) p.loss[b] = ok ? (float)(-(log(zhat) + (double)Ez * 0.6931471805599453)) : INFINITY;
    zinv = ok ? 1.0 / zhat : 0.0;
    if (!ok) return;  // K3 writes zero rows for a sequence without a finite loss
  }
  if (!want_grad) return;
  __syncwarp();

  // ================================================================ phase 2: gamma over p in the tile
  double u[1] = {lane < Lb ? (((Lb - 1 - lane) & 1) ? -1.0 : 1.0) : 0.0};
  int eb = 0;
  double facb = lane == 31 ? 0.0 : 1.0;
  int since = 0;
  // the tile's emissions, checkpoint and lane scale are loaded one tile ahead of their use
  float pe_n[kTT];
  double xa_n;
  int ea_n;
  float* ep2 = e_b + (int64_t)(K - 1) * kTT * estride;  // first emission row of the tile being fetched
  auto fetch = [&](int k, auto full) {
    constexpr bool F = decltype(full)::value;
    const int nrow = F ? kTT : Tb - k * kTT;
    const float* e0 = ep2;
#pragma unroll
    for (int r = 0; r < kTT; ++r) {
      pe_n[r] = ((F || r < nrow) && estore) ? *e0 : 0.f;
      e0 += estride;
    }
    ep2 -= kTT * estride;
    ea_n = 0;
    if (k == 0) {
      xa_n = (lane & 1) ? -1.0 : 1.0;
    } else {
      xa_n = valid ? ckx[k * 32] : 0.0;
      if (k >= 2 && valid) ea_n = cke[(k >> 1) * 32];
    }
  };
  if (Kf < K) fetch(K - 1, std::false_type{});
  else fetch(K - 1, std::true_type{});
  auto tile2 = [&](int k, auto full) {
    constexpr bool F = decltype(full)::value;
    const int nrow = F ? kTT : Tb - k * kTT;
    float* eg = ep2 + (2 * kTT - 1) * estride;  // last row of this tile (ep2: the tile below it, fetched next)
    float pe[kTT][1];
#pragma unroll
    for (int r = 0; r < kTT; ++r) pe[r][0] = pe_n[r];
    double xa[1] = {xa_n};
    const int ea = ea_n;
    if (k > 0) fetch(k - 1, std::true_type{});  // (only a sequence's last tile can be short)
    const int H = ea + eb - Ez;
    const int Ha = max(min(H, 0), -1000);
    const double ga = pow2z(Ha), gb = pow2z(H - Ha) * zinv;
    const int eu = __shfl_up_sync(kFull, ea, 1);
    const double faca = lane == 0 ? 0.0 : pow2z(eu - ea);
    double a[kTT];
#pragma unroll
    for (int r = 0; r < kTT; ++r) {
      if (F || r < nrow) alpha_step<1>(xa, pe[r], faca);
      a[r] = xa[0] * ga;
    }
#pragma unroll
    for (int r = kTT - 1; r >= 0; --r) {
      if (F || r < nrow) {
        double bt[1];
        beta_step<1>(u, bt, pe[r], facb);
        const float g = (float)(a[r] * clamp_big(bt[0] * gb));
        if (estore) *eg = g;
      }
      eg -= estride;
    }
    if (++since == 2 || (!F && nrow < kTT)) {
      since = 0;
      lane_rescale<1, false>(u, eb, facb, lane);
    }
  };
  int kd = K - 1;
  if (Kf < K) tile2(kd--, std::false_type{});
  for (; kd >= 0; --kd) tile2(kd, std::true_type{});
}

template <int NCI>
int launch_bin_seq(const Problem& p, const TiledWs& w, cudaStream_t stream) {
  bin_seq_kernel<NCI><<<(unsigned)p.B, 32, 0, stream>>>(p, w);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

// ------------------------------------------------------------------------------------ K2: lattice on emission tiles
// Register-fed variants of stream::chain_phase1 / chain_phase2 (same arithmetic, same order): the emissions of a full
// tile arrive in registers straight from global memory instead of through a shared-memory p-tile.
template <int NS, int W, int TT>
__device__ __forceinline__ void lat_phase1_regs(double (&x)[NS], stream::ChainScal& c, int lane, double* ck, int* cke, int k,
                                                const double (&pr)[TT][NS]) {
  using namespace stream;
  const int hl = lane & (W - 1);
  double sum[NS];
  if (k > 0) {
    lane_rescale<NS, W, LaneDec<NS, TT>::value>(x, c.e, c.fac, hl);
    if (lane < W) {
#pragma unroll
      for (int j = 0; j < NS; ++j) ck[(k * NS + j) * W] = x[j];
      cke[k * W + lane] = c.e;
    }
  }
#pragma unroll
  for (int i = 0; i < TT; ++i) {
    if (i == 0) chain_step<NS, W, false, true>(x, sum, pr[i], hl, c.carry, c.fac);
    else chain_step<NS, W, false, false>(x, sum, pr[i], hl, 0.0, c.fac);
  }
  c.carry = 0.0;
}
// pr[jj]: the lane's emissions of its jj-th step (alpha walks up the tile, beta down), in the lane's state order
template <int NS, int W, int TT, int AS>
__device__ __forceinline__ void lat_phase2_regs(double (&x)[NS], stream::ChainScal& c, int lane, bool isb, const double (&ckv)[NS],
                                                int eck, int k, const double (&pr)[TT][NS], double* __restrict__ abt) {
  using namespace stream;
  constexpr int Lpad = W * NS;
  const int hl = lane & (W - 1);
  double sum[NS];
  if (!isb) {
    c.e = k == 0 ? 0 : eck;
    c.carry = (k == 0 && hl == 0) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < NS; ++j) x[j] = k == 0 ? 0.0 : ckv[j];
  }
  {
    const int eu = __shfl_up_sync(0xffffffffu, c.e, 1, W);
    if (!isb) c.fac = hl == 0 ? 0.0 : pow2z(eu - c.e);
  }
  // beta entry = beta * 2^H / Zhat, H = e_beta + e_alpha - E_z.  H can fall below the float64 exponent range while the
  // product with the alpha entry is of order one (alpha and beta both large in their lanes' scales: mass that has just
  // arrived through lanes scaled for it -- every state of a sequence with a single admissible path and emissions near
  // one), so the power of two is applied in two exact steps
  const int Hs = isb ? c.e + (k == 0 ? 0 : eck) - c.Ez : 0;
  const double bs = isb ? pow2z(max(Hs, -1000)) * c.zinv : 1.0;
  const double bs2 = pow2z(min(Hs + 1000, 0));
  const int s0 = isb ? (Lpad - 1 - hl * NS) : hl * NS;
  const int sdir = isb ? -1 : 1;
  double* dst = abt + (isb ? TT * AS : 0) + s0;
#pragma unroll
  for (int jj = 0; jj < TT; ++jj) {
    const int i = isb ? (TT - 1 - jj) : jj;
    if (jj == 0) chain_step<NS, W, true, true>(x, sum, pr[jj], hl, c.carry, c.fac);
    else chain_step<NS, W, true, false>(x, sum, pr[jj], hl, 0.0, c.fac);
#pragma unroll
    for (int j = 0; j < NS; ++j) dst[i * AS + sdir * j] = isb ? fmin(sum[j] * bs, 1e300) * bs2 : x[j];
  }
  c.carry = 0.0;
  int e2 = c.e;
  double f2 = c.fac;
  lane_rescale<NS, W, LaneDec<NS, TT>::value>(x, e2, f2, hl);
  if (isb) {
    c.e = e2;
    c.fac = f2;
  }
}

// The emission tile holds p_t(s) = exp(e_t(s) - rowc_t) (K1; zeros for states >= L_b).  NS = 2 and an even Lmax: full
// tiles go through registers (each lane loads the float pairs of its own two states, one tile ahead); everything else
// is staged through the warp's shared-memory p-tile.
// NS = 2: at most 72 registers, so that 28 warps share an SM (4096 sequences are then a single wave on 148 SMs)
template <int NS>
__global__ void __launch_bounds__(kLatWarps * 32, NS == 2 ? 7 : 1) lattice_tile_kernel(Problem p, TiledWs w) {
  if (*w.flag != 0) return;
  using namespace stream;
  constexpr int W = 16, TT = 8, Lpad = 16 * NS, PS = Lpad + 8, AS = Lpad + 8;
  constexpr int kWarpBytes = TT * PS * 4 + 2 * TT * AS * 8;
  extern __shared__ __align__(16) unsigned char smraw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * kLatWarps + warp;
  if (b >= p.B) return;
  float* pt = reinterpret_cast<float*>(smraw + (size_t)warp * kWarpBytes);
  double* abt = reinterpret_cast<double*>(pt + TT * PS);
  const int64_t Tb64 = p.in_len[b], Lb64 = p.tgt_len[b];
  if (!seq_feasible(Tb64, Lb64, p.T, p.Lmax)) {
    if (lane == 0) p.loss[b] = INFINITY;
    return;
  }
  const int Tb = (int)Tb64, Lb = (int)Lb64, Lmax = (int)p.Lmax;
  const int NTb = (Tb + TT - 1) / TT;
  double* ck = w.ckpt + ((size_t)b * w.NT) * Lpad + (lane & (W - 1));
  int* cke = w.cke + (size_t)b * w.NT * W;  // [NT][W] lane scales of the checkpoints
  const bool isb = lane >= 16;

  constexpr int NPR = Lpad / 32;  // floats per lane and tile row
  const int64_t estride = p.B * (int64_t)Lmax;
  float* const e_b = w.emis + b * Lmax + lane;
  bool sv[NPR];
#pragma unroll
  for (int q = 0; q < NPR; ++q) sv[q] = lane + 32 * q < Lb;
  // tile k -> shared-memory p-tile (any tile; zeros beyond T_b and L_b)
  auto stage = [&](int k) {
    const int t0 = k * TT, nv = min(TT, Tb - t0);
    const float* pe = e_b + (int64_t)t0 * estride;
#pragma unroll
    for (int r = 0; r < TT; ++r) {
#pragma unroll
      for (int q = 0; q < NPR; ++q) pt[r * PS + lane + 32 * q] = (r < nv && sv[q]) ? __ldg(pe + 32 * q) : 0.f;
      pe += estride;
    }
  };
  // register path (NS = 2): the lane's state pair, forward (phase 1, alpha lanes of phase 2) or reversed (beta lanes)
  const bool regs = NS == 2 && (Lmax & 1) == 0;
  const int NTf = regs ? Tb / TT : 0;  // tiles [0, NTf) are full and take the register path
  float2 pv[TT];
  auto fetch2 = [&](int k, bool phase2) {
    const bool rev = phase2 && isb;
    const int s = rev ? (W - 1 - (lane & (W - 1))) * 2 : (lane & (W - 1)) * 2;
    const bool okp = s < Lmax;  // the pair is inside the row (Lmax is even); states >= L_b hold zeros
    const float* pe = w.emis + (((int64_t)k * TT + (rev ? TT - 1 : 0)) * p.B + b) * Lmax + s;
    const int64_t step = rev ? -estride : estride;
#pragma unroll
    for (int r = 0; r < TT; ++r) {
      pv[r] = okp ? __ldg(reinterpret_cast<const float2*>(pe)) : make_float2(0.f, 0.f);
      pe += step;
    }
  };
  auto widen = [&](double (&pr)[TT][2], bool phase2) {
    const bool rev = phase2 && isb;
#pragma unroll
    for (int r = 0; r < TT; ++r) {
      pr[r][0] = (double)(rev ? pv[r].y : pv[r].x);
      pr[r][1] = (double)(rev ? pv[r].x : pv[r].y);
    }
  };

  ChainScal chain;
  double cx[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) cx[j] = 0.0;
  chain.carry = (lane == 0) ? 1.0 : 0.0;
  chain.zinv = 0.0;
  chain.e = 0; chain.Ez = 0;
  chain.fac = (lane & (W - 1)) == 0 ? 0.0 : 1.0;
  // ---- phase 1: alpha, one checkpoint per tile
  if (NTf > 0) fetch2(0, false);
  for (int k = 0; k < NTb; ++k) {
    if constexpr (NS == 2) {
      if (k < NTf) {
        double pr[TT][2];
        widen(pr, false);
        if (k + 1 < NTf) fetch2(k + 1, false);
        lat_phase1_regs<2, W, TT>(cx, chain, lane, ck, cke, k, pr);
        continue;
      }
    }
    stage(k);
    __syncwarp();
    chain_phase1<NS, W, TT, PS>(cx, chain, lane, Tb, ck, cke, k, pt);
    __syncwarp();
  }
  chain_readout<NS, W>(cx, chain, lane, Lb, &p.loss[b], 1.f, nullptr);
  if (p.grad == nullptr) return;
  // ---- phase 2: tiles downwards; lanes 0-15 replay alpha from the checkpoint, lanes 16-31 run beta
  if (NTf > 0) fetch2(NTf - 1, true);
  for (int k = NTb - 1; k >= 0; --k) {
    double ckv[NS];
    int eck = 0;
    if (k > 0) {
#pragma unroll
      for (int j = 0; j < NS; ++j) ckv[j] = ck[(k * NS + j) * W];
      // the alpha direction's own lane scale; the beta direction takes the scale of the alpha lane with the same states
      eck = cke[k * W + (isb ? W - 1 - (lane & (W - 1)) : (lane & (W - 1)))];
    } else {
#pragma unroll
      for (int j = 0; j < NS; ++j) ckv[j] = 0.0;
    }
    bool done = false;
    if constexpr (NS == 2) {
      if (k < NTf) {
        double pr[TT][2];
        widen(pr, true);
        if (k > 0) fetch2(k - 1, true);
        lat_phase2_regs<2, W, TT, AS>(cx, chain, lane, isb, ckv, eck, k, pr, abt);
        done = true;
      }
    }
    if (!done) {
      stage(k);
      __syncwarp();
      chain_phase2<NS, W, TT, PS, AS>(cx, chain, lane, isb, Tb, ckv, eck, k, pt, abt);
    }
    __syncwarp();
    // gamma_t(s) = alpha entry * beta entry (chain_phase2 folds 1/Z and the lane scales into the beta tile) -> emission tile
    const int t0 = k * TT, nv = min(TT, Tb - t0);
    float* pe = e_b + (int64_t)t0 * estride;
#pragma unroll
    for (int r = 0; r < TT; ++r) {
#pragma unroll
      for (int q = 0; q < NPR; ++q) {
        const int st = lane + 32 * q;
        if (r < nv && sv[q]) pe[32 * q] = (float)(abt[r * AS + st] * abt[(TT + r) * AS + st]);
      }
      pe += estride;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------ row staging by TMA
// A warp's batch = kTB rows (t..t+kTB-1, b) of the logits: each row is C floats, only 4-byte aligned when C % 4 != 0,
// and consecutive rows are B*C floats apart.  Lanes 0..nrow-1 each move the 16-byte aligned superset of one row
// with one cp.async.bulk into the row's slot; lane 0 arrives on the buffer's mbarrier with the byte total.  Two
// buffers per warp: the next batch lands while the current one is worked on, at no register cost.
__device__ __forceinline__ void bulk_g2s_plain(uint32_t dst_smem, uint64_t src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}
template <int TB>
struct RowStage {
  unsigned char* buf;  // [2][TB][RSB] this warp's slots
  uint64_t* bar;       // [2] this warp's mbarriers
  uint32_t par;        // bit s = parity of the next completion of buffer s
  int RSB;             // slot bytes = align16(4*C + 12)
  uint64_t lim;        // end of the logits tensor
  int64_t rstride;     // floats between consecutive rows of a sequence
  int C, lane;

  __device__ __forceinline__ void init(unsigned char* buf_, uint64_t* bar_, const Problem& p, int lane_) {
    buf = buf_; bar = bar_; par = 0u; lane = lane_;
    C = (int)p.C;
    RSB = (4 * C + 12 + 15) & ~15;
    rstride = p.B * p.C;
    lim = reinterpret_cast<uint64_t>(p.logits) + (uint64_t)p.T * p.B * p.C * 4u;
    if (lane < 2) stream::mbar_init(&bar[lane], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
  }
  // rows x0 + r*rstride, r < nrow -> buffer s (every lane of the warp calls this)
  __device__ __forceinline__ void issue(int s, const float* x0, int nrow) {
    uint32_t bytes = 0;
    if (lane < nrow) {
      const uint64_t a = reinterpret_cast<uint64_t>(x0 + lane * rstride);
      const uint64_t a0 = a & ~uint64_t(15);
      uint64_t a1 = (a + 4u * C + 15) & ~uint64_t(15);
      unsigned char* dst = buf + (size_t)(s * TB + lane) * RSB;
      if (a1 > lim) {
        // the tensor's last row ends inside a 16-byte chunk: the bulk copy stops before it, the rest goes by hand
        a1 = lim & ~uint64_t(15);
        for (uint64_t q = max(a1, a); q < a + 4u * C; q += 4)
          *reinterpret_cast<float*>(dst + (q - a0)) = __ldg(reinterpret_cast<const float*>(q));
      }
      if (a1 > a0) {
        bytes = (uint32_t)(a1 - a0);
        bulk_g2s_plain(stream::smem_u32(dst), a0, bytes, stream::smem_u32(&bar[s]));
      }
    }
    const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);  // also orders the by-hand stores before the arrival
    if (lane == 0) {
      if (total) stream::mbar_arrive_expect_tx(&bar[s], total);  // the phase cannot complete before this arrival
      else stream::mbar_arrive(&bar[s]);
    }
  }
  __device__ __forceinline__ void wait(int s) {
    stream::mbar_wait(&bar[s], (par >> s) & 1u);
    par ^= 1u << s;
  }
  // row r of buffer s, first float of the row (rows keep their 16-byte phase)
  __device__ __forceinline__ const float* row(int s, int r, const float* x0) const {
    const uint32_t ph = (uint32_t)(reinterpret_cast<uint64_t>(x0 + r * rstride) & 15u);
    return reinterpret_cast<const float*>(buf + (size_t)(s * TB + r) * RSB + ph);
  }
};

// ------------------------------------------------------------------------------------ K3: gradient
// Batches of kGTB = 4 rows and 16 warps per CTA: the staging buffers of a warp are half as large, so that twice as
// many warps share an SM (the kernel is issue-bound at 16 warps per SM).
constexpr int kGTB = 4;
constexpr int kGradWarps = 16;
template <int NCI, int Lp>
__global__ void __launch_bounds__(kGradWarps * 32, 2) bin_grad_kernel(Problem p, TiledWs w) {
  if (*w.flag != 0) return;
  constexpr int kTB = kGTB, kRowWarps = kGradWarps;  // this kernel's batch geometry
  extern __shared__ __align__(128) unsigned char smg[];  // [kRowWarps][2][kTB][RSB] logits rows | [kRowWarps][kTB][Lp] gamma
  __shared__ uint64_t bars[kRowWarps][2];
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
  RowStage<kTB> rs;
  const int RSB = (4 * C + 12 + 15) & ~15;
  rs.init(smg + (size_t)warp * 2 * kTB * RSB, bars[warp], p, lane);
  // gamma tile of the warp: [kTB][Lp + 1]; column Lp stays zero (what a lane with no state left adds)
  constexpr int LS = Lp + 1;
  float* gs_all = reinterpret_cast<float*>(smg + (size_t)kRowWarps * 2 * kTB * RSB);
  float* gs = gs_all + (size_t)warp * kTB * LS;
  if (lane < kTB) gs[lane * LS + Lp] = 0.f;
  const int64_t rstride = rs.rstride;
  const int64_t gstride = p.B * (int64_t)Lmax;
  const uint32_t* cm = w.cmask + (size_t)b * C * LW;
  const float* const gam_b = w.emis + b * Lmax + lane;
  const bool lane_state = lane < Lb;
  // requests run one batch ahead: the next batch's logits (TMA) and gamma rows (registers) are in flight while the
  // current one is walked and stored
  float gn[kTB];
  auto request_gamma = [&](int64_t tb0, int nrow) {
    const float* g = gam_b + tb0 * gstride;
#pragma unroll
    for (int r = 0; r < kTB; ++r) {
      gn[r] = (r < nrow && lane_state) ? __ldg(g) : 0.f;
      g += gstride;
    }
  };
  auto x_of = [&](int64_t tb0) { return p.logits + (tb0 * p.B + b) * p.C; };
  auto rows_of = [&](int64_t tb0) { return (int)min((int64_t)kTB, tlive - tb0); };
  const int64_t tfirst = t0 + (int64_t)warp * kTB;
  if (tfirst < tlive) {
    rs.issue(0, x_of(tfirst), rows_of(tfirst));
    request_gamma(tfirst, rows_of(tfirst));
  }
  int sb = 0;
  // Lmax <= 32: the states of every class, decoded once per CTA into a shared-memory table rt[round][class] of
  // gamma-tile columns (ascending state order; used-up rounds point at the zero column); rreg = the warp's rounds
  constexpr int kCs = 32 * NCI;
  unsigned char* rt = reinterpret_cast<unsigned char*>(gs_all + (size_t)kRowWarps * kTB * LS);
  int rreg[NCI];
#pragma unroll
  for (int i = 0; i < NCI; ++i) {
    const int c = lane + 32 * i;
    const uint32_t m = (LW == 1 && (i + 1 < NCI || c < C)) ? __ldg(cm + c) : 0u;
    rreg[i] = __reduce_max_sync(0xffffffffu, __popc(m));
  }
  if (LW == 1) {
    for (int c = threadIdx.x; c < kCs; c += kRowWarps * 32) {
      uint32_t m = c < C ? __ldg(cm + c) : 0u;
      for (int k = 0; k < 32; ++k) {
        rt[k * kCs + c] = (unsigned char)(m ? __ffs(m) - 1 : Lp);
        m &= m - 1u;
      }
    }
    __syncthreads();  // (every warp of the CTA gets here: the exits above do not depend on the warp)
  }
  // one batch; kFull: all kTB rows are live (no per-row predicates)
  auto batch = [&](int64_t tb0, auto full) {
    constexpr bool kFull = decltype(full)::value;
    const int nrow = kFull ? kTB : rows_of(tb0);
    const int64_t tnext = tb0 + kRowWarps * kTB;
    if (tnext < tlive) rs.issue(sb ^ 1, x_of(tnext), rows_of(tnext));
#pragma unroll
    for (int r = 0; r < kTB; ++r) gs[r * LS + lane] = gn[r];
    if (Lb > 32) {  // states beyond the first 32 are not prefetched
      const float* gam = w.emis + (tb0 * p.B + b) * Lmax;
      for (int s = lane + 32; s < Lb; s += 32) {
#pragma unroll
        for (int r = 0; r < kTB; ++r) gs[r * LS + s] = r < nrow ? __ldg(gam + r * gstride + s) : 0.f;
      }
    }
    if (tnext < tlive) request_gamma(tnext, rows_of(tnext));
    rs.wait(sb);
    const float* x0 = x_of(tb0);
    float acc[NCI][kTB];
#pragma unroll
    for (int r = 0; r < kTB; ++r) {
      const float* xr = rs.row(sb, r, x0) + lane;
#pragma unroll
      for (int i = 0; i < NCI; ++i) {
        const float v = ((i + 1 < NCI || lane + 32 * i < C) && (kFull || r < nrow)) ? xr[32 * i] : 0.f;
        acc[i][r] = sigmoid_fast(v);
      }
    }
    __syncwarp();
    // sum_s gamma_t(s) over the states of the lane's classes: the warp runs as many rounds as its fullest mask
    // needs (uniform trip count); a lane whose mask is used up reads the zero column
#pragma unroll
    for (int i = 0; i < NCI; ++i) {
      const int c = lane + 32 * i;
      if (LW == 1) {
        const unsigned char* rc = rt + c;
        for (int k = 0; k < rreg[i]; ++k) {
          const float* col = gs + rc[k * kCs];
#pragma unroll
          for (int r = 0; r < kTB; ++r) acc[i][r] -= col[r * LS];
        }
      } else {
        for (int wd = 0; wd < LW; ++wd) {
          uint32_t m = (i + 1 < NCI || c < C) ? __ldg(cm + (size_t)c * LW + wd) : 0u;
          const int rounds = __reduce_max_sync(0xffffffffu, __popc(m));
          const float* gw = gs + wd * 32;
          for (int k = 0; k < rounds; ++k) {
            const float* col = m ? gw + (__ffs(m) - 1) : gs + Lp;
            m &= m - 1u;
#pragma unroll
            for (int r = 0; r < kTB; ++r) acc[i][r] -= col[r * LS];
          }
        }
      }
    }
    float* g0 = p.grad + (tb0 * p.B + b) * p.C + lane;
#pragma unroll
    for (int r = 0; r < kTB; ++r) {
      if (kFull || r < nrow) {
#pragma unroll
        for (int i = 0; i < NCI; ++i)
          if (i + 1 < NCI || lane + 32 * i < C) g0[32 * i] = wC * acc[i][r];
      }
      g0 += rstride;
    }
    __syncwarp();  // every lane is done with buffer sb and the gamma tile before they are refilled
  };
  for (int64_t tb0 = tfirst; tb0 < tlive; tb0 += kRowWarps * kTB, sb ^= 1) {
    if (tb0 + kTB <= tlive) batch(tb0, std::true_type{});
    else batch(tb0, std::false_type{});
  }
}

template <int NS>
int launch_lattice(const Problem& p, const TiledWs& w, cudaStream_t stream) {
  constexpr int Lpad = 16 * NS, PS = Lpad + 8, AS = Lpad + 8;
  constexpr int kWarpBytes = 8 * PS * 4 + 2 * 8 * AS * 8;
  const size_t smem = (size_t)kLatWarps * kWarpBytes;
  auto kern = lattice_tile_kernel<NS>;
  if (smem > 48 * 1024) NBCTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)((p.B + kLatWarps - 1) / kLatWarps), kLatWarps * 32, smem, stream>>>(p, w);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

template <int NCI, int Lp>
int launch_grad_inst(const Problem& p, const TiledWs& w, dim3 grid, cudaStream_t stream) {
  const size_t RSB = (4 * (size_t)p.C + 12 + 15) & ~(size_t)15;
  const size_t smem = (size_t)kGradWarps * 2 * kGTB * RSB + sizeof(float) * kGradWarps * kGTB * (Lp + 1) +
                      (Lp == 32 ? (size_t)32 * 32 * NCI : 0);  // + the round table (Lmax <= 32)
  auto kern = bin_grad_kernel<NCI, Lp>;
  if (smem > 48 * 1024) NBCTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kGradWarps * 32, smem, stream>>>(p, w);
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
  const size_t smem = sizeof(float) * kRowWarps * kTB * (32 * NCI + 8) + (size_t)p.Lmax * 32;
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
  w.cmask = reinterpret_cast<uint32_t*>(c + l.o_cmask);
  w.ckpt = reinterpret_cast<double*>(c + l.o_ckpt);
  w.cke = reinterpret_cast<int*>(c + l.o_cke);
  w.emis = reinterpret_cast<float*>(c + l.o_emis);
  w.LW = l.LW; w.NT = l.NT; w.Lpad = l.Lpad;
  w.rec = c + l.o_rec; w.rec_bytes = (int64_t)l.rec_bytes; w.o_rec_ckx = l.o_rec_ckx;
  w.rs32 = (int)std::min<int64_t>(p.B * p.C, INT32_MAX); w.es32 = (int)std::min<int64_t>(p.B * p.Lmax, INT32_MAX);
  *flag_out = w.flag;
  // the flag and the class masks (set with atomicOr by the emission kernel) are contiguous: one memset
  NBCTC_CUDA_CHECK(cudaMemsetAsync(w.flag, 0, l.o_cmask + sizeof(uint32_t) * (size_t)p.B * p.C * l.LW, stream));
  const dim3 grid((unsigned)p.B, (unsigned)((p.T + kTCh - 1) / kTCh));
  int rc;
  const int nci = (int)((p.C + 31) / 32);
  static const bool no_seq = getenv("NBCTC_BIN_NOSEQ") != nullptr;
  if (l.seq && !no_seq) {  // emissions + lattice in one sequence-per-warp kernel
    switch (nci) {
      case 1: rc = launch_bin_seq<1>(p, w, stream); break;
      case 2: rc = launch_bin_seq<2>(p, w, stream); break;
      case 3: rc = launch_bin_seq<3>(p, w, stream); break;
      case 4: rc = launch_bin_seq<4>(p, w, stream); break;
      case 5: rc = launch_bin_seq<5>(p, w, stream); break;
      case 6: rc = launch_bin_seq<6>(p, w, stream); break;
      case 7: rc = launch_bin_seq<7>(p, w, stream); break;
      default: rc = launch_bin_seq<8>(p, w, stream); break;
    }
    if (rc != NBCTC_OK || p.grad == nullptr) return rc;
    goto gradient;
  }
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

gradient:
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
