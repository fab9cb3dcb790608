// Fused no-blank CTC forward+backward kernel for sm_100a (see nbctc_fused.cu for the overview).
//
// One CTA per sequence b, warp-specialised:
//   warp 0      "chain" warp: lattice recursions, lane = NS consecutive states (NoBlankCTC.py:71-87)
//   warps 1..NW "row" warps : stream (t,b) rows of the logits, LPR lanes per row, CPL 16-byte chunks per lane
// Template parameters fix the row geometry at compile time so the streaming code is branch-light:
//   NS  states per chain lane (1,2,4,8  -> Lmax <= 32*NS)
//   LPR lanes per row (4 or 8), R = 32/LPR rows per warp pass, kTT/R passes per tile
//   CPL chunks per lane per row segment
#pragma once

#include <cuda_runtime.h>

#include "common.cuh"

namespace nbctc {

constexpr int kTT = 8;        // time steps per tile
constexpr int kMaxBuf = 8;    // p-/gamma-tile ring depth upper bound
constexpr int kMaxSlot = 16;  // row-tile (TMA) ring depth upper bound
constexpr int kNW = 4;        // row warps per CTA
constexpr int kThreads = 32 * (kNW + 2);  // chain warp + row warps + TMA producer warp

struct FusedCfg {
  int NS, Lpad;
  int LPR, CPL, NSEG;
  int NBUFP, NBUFG; // p-tile / gamma-tile ring depths
  int NSLOT;        // row-tile ring depth (TMA bulk copies land here)
  int RS;           // bytes per row slot = 16 * max chunks per row
  int NTmax;        // ceil(T / kTT)
  int Cd;           // floats per scatter buffer
  int ckpt_global;  // checkpoints live in the workspace instead of shared memory
  int lse_global;
  uint32_t o_bar, o_lab, o_lse, o_ckpt, o_cke, o_ptile, o_gtile, o_atile, o_delta, o_ring, smem_bytes;
  double* ws_ckpt;  // [B][NTmax][Lpad]
  int* ws_cke;      // [B][NTmax]
  float* ws_lse;    // [B][T]
  const float* logits_end;  // one past the last logit (bulk copies never read past its 16-byte round-up)
};

int launch_fused_ns1(const Problem& p, const FusedCfg& cfg, cudaStream_t stream);
int launch_fused_ns2(const Problem& p, const FusedCfg& cfg, cudaStream_t stream);
int launch_fused_ns4(const Problem& p, const FusedCfg& cfg, cudaStream_t stream);
int launch_fused_ns8(const Problem& p, const FusedCfg& cfg, cudaStream_t stream);

#ifdef __CUDACC__
namespace fused {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kPMin = 7.52316385e-37f;  // 2^-120: emission floor (8 steps stay inside the f64 range)
constexpr float kNegInf = -INFINITY;

// ---------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (1-D, 16-byte aligned, size multiple of 16), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "NBCTC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra NBCTC_DONE;\n"
      "bra NBCTC_WAIT;\n"
      "NBCTC_DONE:\n"
      "}\n" ::"r"(smem_u32(b)),
      "r"(parity), "r"(2000000u)
      : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_f4_hint(const float4* ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg_f4_hint(float4* ptr, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void stg_f_hint(float* ptr, float v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(ptr), "f"(v), "l"(pol) : "memory");
}

template <int NS>
__device__ __forceinline__ int rescale_pow2(double (&v)[NS]) {
  double m = v[0];
#pragma unroll
  for (int j = 1; j < NS; ++j) m = fmax(m, v[j]);
  unsigned hi = (unsigned)__double2hiint(m);  // values are >= 0
  unsigned mx = __reduce_max_sync(0xffffffffu, hi);
  int ex = (int)(mx >> 20);
  if (ex == 0 || ex >= 0x7ff) return 0;
  int e = ex - 1023;
  double sc = __hiloint2double((1023 - e) << 20, 0);  // exact 2^-e
#pragma unroll
  for (int j = 0; j < NS; ++j) v[j] *= sc;
  return e;
}
__device__ __forceinline__ double pow2i(int e) {  // exact 2^e, e clamped to the normal range
  e = max(-1022, min(1023, e));
  return __hiloint2double((1023 + e) << 20, 0);
}

struct Smem {
  uint64_t *pfull, *pempty, *gfull, *gempty, *sfull, *sempty;
  unsigned char* ring;
  int* lab;
  float* lse;
  double* ckpt;
  int* cke;
  float* ptile;
  float* gtile;
  double* atile;
  float* delta;
};

// ============================================================================ chain warp
template <int NS>
__device__ __forceinline__ void load_p(const float* src, double (&p)[NS]) {
  if constexpr (NS == 1) {
    p[0] = (double)src[0];
  } else if constexpr (NS == 2) {
    float2 v = *reinterpret_cast<const float2*>(src);
    p[0] = v.x; p[1] = v.y;
  } else {
#pragma unroll
    for (int j = 0; j < NS; j += 4) {
      float4 v = *reinterpret_cast<const float4*>(src + j);
      p[j] = v.x; p[j + 1] = v.y; p[j + 2] = v.z; p[j + 3] = v.w;
    }
  }
}
template <int NS>
__device__ __forceinline__ void store_g(float* dst, const float (&g)[NS]) {
  if constexpr (NS == 1) {
    dst[0] = g[0];
  } else if constexpr (NS == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(g[0], g[1]);
  } else {
#pragma unroll
    for (int j = 0; j < NS; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
  }
}

// alpha_t(s) = (alpha_{t-1}(s) + alpha_{t-1}(s-1)) * p_t(s); `carry` enters state 0 (1.0 only at t = 0: the
// virtual start state, NoBlankCTC.py:92-93 forward_prob[:,0] = 0 and the t>0 shift guard :75)
template <int NS>
__device__ __forceinline__ void alpha_step(double (&a)[NS], const double (&p)[NS], int lane, double& carry) {
  double up = __shfl_up_sync(0xffffffffu, a[NS - 1], 1);
  if (lane == 0) up = carry;
  carry = 0.0;
#pragma unroll
  for (int j = NS - 1; j >= 1; --j) a[j] = (a[j] + a[j - 1]) * p[j];
  a[0] = (a[0] + up) * p[0];
}

template <int NS>
__device__ __forceinline__ void chain_warp(const Problem& P, const FusedCfg& cfg, const Smem& S, int lane, int64_t b,
                                           int Tb, int Lb, float wgt) {
  constexpr int Lpad = 32 * NS;
  constexpr bool kRegTile = NS <= 2;  // tile-local alpha replay in registers
  const int NT = (Tb + kTT - 1) / kTT;
  const int NBUF = cfg.NBUFP, NBUFG = cfg.NBUFG;
  double* ck = (cfg.ckpt_global ? cfg.ws_ckpt + ((size_t)b * cfg.NTmax) * Lpad : S.ckpt) + lane;
  int* cke = cfg.ckpt_global ? cfg.ws_cke + (size_t)b * cfg.NTmax : S.cke;
  double a[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) a[j] = 0.0;
  int Ea = 0;
  double carry = 1.0;
  // ------------------------------------------------------------------ phase 1: alpha
  for (int k = 0; k < NT; ++k) {
    const int buf = k % NBUF;
    if (k > 0) {
      Ea += rescale_pow2<NS>(a);
#pragma unroll
      for (int j = 0; j < NS; ++j) ck[(k * NS + j) * 32] = a[j];
      if (lane == 0) cke[k] = Ea;
    }
    mbar_wait(&S.pfull[buf], (k / NBUF) & 1);
    const float* pt = S.ptile + buf * (kTT * Lpad) + lane * NS;
    const int nv = min(kTT, Tb - k * kTT);
    if (nv == kTT) {
#pragma unroll
      for (int i = 0; i < kTT; ++i) {
        double p[NS];
        load_p<NS>(pt + i * Lpad, p);
        alpha_step<NS>(a, p, lane, carry);
      }
    } else {
      for (int i = 0; i < nv; ++i) {
        double p[NS];
        load_p<NS>(pt + i * Lpad, p);
        alpha_step<NS>(a, p, lane, carry);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.pempty[buf]);
  }
  // ------------------------------------------------------------------ read-out (NoBlankCTC.py:58-68,:139)
  const int sl = Lb - 1;
  double mine = 0.0;
#pragma unroll
  for (int j = 0; j < NS; ++j)
    if (j == sl % NS) mine = a[j];
  const double zhat = __shfl_sync(0xffffffffu, mine, sl / NS);
  const int Ez = Ea;
  if (lane == 0) P.loss[b] = (zhat > 0.0) ? (float)(-(log(zhat) + (double)Ez * 0.6931471805599453)) : INFINITY;
  if (P.grad == nullptr) return;
  const double zinv = (zhat > 0.0) ? (double)wgt / zhat : 0.0;  // sequence weight folded into gamma
  // ------------------------------------------------------------------ phase 2: beta, gamma
  // u(s) = beta_{t+1}(s) p_{t+1}(s).  Virtual start: u_{T_b}(L_b) = 1 makes beta_{T_b-1}(L_b-1) = 1 with no branch
  // (state L_b itself has p = 0 and alpha = 0, so it contributes nothing).
  double u[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) u[j] = (lane * NS + j == Lb) ? 1.0 : 0.0;
  double bcarry = (Lb == Lpad) ? 1.0 : 0.0;
  int Eb = 0;
  for (int j2 = 0; j2 < NT; ++j2) {
    const int k = NT - 1 - j2;
    const int n = NT + j2;
    const int buf = n % NBUF;
    const int gbuf = j2 % NBUFG;
    int EaK = 0;
    if (k == 0) {
#pragma unroll
      for (int j = 0; j < NS; ++j) a[j] = 0.0;
      carry = 1.0;
    } else {
#pragma unroll
      for (int j = 0; j < NS; ++j) a[j] = ck[(k * NS + j) * 32];
      EaK = cke[k];
    }
    // gamma = alpha * beta * w / Z: the power-of-two part is split over both factors (range safety)
    const int d = EaK + Eb - Ez;
    const double s1 = pow2i(d / 2);
    const double s2 = pow2i(d - d / 2) * zinv;
#pragma unroll
    for (int j = 0; j < NS; ++j) a[j] *= s1;  // exact; alpha replay runs pre-scaled
    if (k == 0) carry = s1;
    mbar_wait(&S.pfull[buf], (n / NBUF) & 1);
    const float* pt = S.ptile + buf * (kTT * Lpad) + lane * NS;
    const int nv = min(kTT, Tb - k * kTT);
    double ar[kRegTile ? kTT : 1][NS];
    double* at = S.atile + lane;
    // replay alpha inside the tile
#pragma unroll
    for (int i = 0; i < kTT; ++i) {
      if (i < nv) {
        double p[NS];
        load_p<NS>(pt + i * Lpad, p);
        alpha_step<NS>(a, p, lane, carry);
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          if constexpr (kRegTile) ar[i][j] = a[j];
          else at[(i * NS + j) * 32] = a[j];
        }
      }
    }
    if (j2 >= NBUFG) mbar_wait(&S.gempty[gbuf], ((j2 / NBUFG) - 1) & 1);
    float* gt = S.gtile + gbuf * (kTT * Lpad) + lane * NS;
#pragma unroll
    for (int ii = 0; ii < kTT; ++ii) {
      const int i = kTT - 1 - ii;
      if (i < nv) {
        double p[NS], beta[NS];
        load_p<NS>(pt + i * Lpad, p);
        double dn = __shfl_down_sync(0xffffffffu, u[0], 1);
        if (lane == 31) dn = bcarry;
        bcarry = 0.0;
#pragma unroll
        for (int j = 0; j < NS - 1; ++j) beta[j] = u[j] + u[j + 1];
        beta[NS - 1] = u[NS - 1] + dn;
        float g[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          double al;
          if constexpr (kRegTile) al = ar[i][j];
          else al = at[(i * NS + j) * 32];
          g[j] = (float)(al * (beta[j] * s2));
          u[j] = beta[j] * p[j];
        }
        store_g<NS>(gt + i * Lpad, g);
      }
    }
    Eb += rescale_pow2<NS>(u);
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&S.gfull[gbuf]);
      mbar_arrive(&S.pempty[buf]);
    }
  }
}

// ============================================================================ TMA producer warp
// One elected lane walks the sequence's rows tile by tile -- phase 1 upwards, phase 2 downwards (most recently
// read rows first, so the re-read hits L2) -- and issues one bulk copy per row (the 16-byte aligned superset
// of the 4-byte aligned row) into the shared-memory ring.  Phase-1 copies carry an L2 evict_last policy, phase-2
// copies evict_first.
__device__ __forceinline__ void producer_warp(const Problem& P, const FusedCfg& cfg, const Smem& S, int lane, int64_t b,
                                              int Tb) {
  if (lane != 0) return;
  const int NT = (Tb + kTT - 1) / kTT;
  const int total = (P.grad != nullptr) ? 2 * NT : NT;
  const int C = (int)P.C;
  const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
  const uintptr_t safe_end = (reinterpret_cast<uintptr_t>(cfg.logits_end) + 15) & ~uintptr_t(15);
  const bool end_unaligned = (reinterpret_cast<uintptr_t>(cfg.logits_end) & 15) != 0;
  for (int n = 0; n < total; ++n) {
    const int k = n < NT ? n : 2 * NT - 1 - n;
    const int slot = n % cfg.NSLOT;
    const int use = n / cfg.NSLOT;
    if (use >= 1) mbar_wait(&S.sempty[slot], (use - 1) & 1);
    const int nv = min(kTT, Tb - k * kTT);
    unsigned char* dst = S.ring + (size_t)slot * kTT * cfg.RS;
    const float* row0 = P.logits + ((int64_t)k * kTT * P.B + b) * C;
    const int64_t strideT = P.B * (int64_t)C;
    uint32_t bytes = 0;
    for (int i = 0; i < nv; ++i) {
      const uintptr_t a = reinterpret_cast<uintptr_t>(row0 + i * strideT);
      const uintptr_t a0 = a & ~uintptr_t(15);
      const uint32_t nb = (uint32_t)(((a + (uintptr_t)C * 4 + 15) & ~uintptr_t(15)) - a0);
      // the very last row of an unaligned tensor would be over-read by < 16 bytes: copy it by hand instead
      if (end_unaligned && a0 + nb == safe_end) {
        const float* src = reinterpret_cast<const float*>(a);
        float* d = reinterpret_cast<float*>(dst + (size_t)i * cfg.RS) + ((a >> 2) & 3);
        for (int c = 0; c < C; ++c) d[c] = __ldg(src + c);
      } else {
        bytes += nb;
      }
    }
    mbar_arrive_expect_tx(&S.sfull[slot], bytes);
    const uint64_t pol = n < NT ? pol_keep : pol_stream;
    for (int i = 0; i < nv; ++i) {
      const uintptr_t a = reinterpret_cast<uintptr_t>(row0 + i * strideT);
      const uintptr_t a0 = a & ~uintptr_t(15);
      const uint32_t nb = (uint32_t)(((a + (uintptr_t)C * 4 + 15) & ~uintptr_t(15)) - a0);
      if (!(end_unaligned && a0 + nb == safe_end))
        bulk_g2s(dst + (size_t)i * cfg.RS, reinterpret_cast<const void*>(a0), nb, &S.sfull[slot], pol);
    }
  }
}

// ============================================================================ row warps
// Geometry of one (t,b) row seen as 16-byte chunks: the row starts `off4` floats into chunk 0 and
// ends `rem` floats into chunk nch-1 (rows are only 4-byte aligned when C % 4 != 0).  `srow` is the row's
// copy in the shared-memory ring (same 16-byte phase as in global memory).
struct RowGeom {
  const float4* srow;
  int64_t goff;        // element offset of the row in logits / grad
  int off4, nch, rem;  // rem in 1..4 = valid floats in the last chunk
};
__device__ __forceinline__ void mask_head(float4& v, int off4) {
  if (off4 > 0) v.x = kNegInf;
  if (off4 > 1) v.y = kNegInf;
  if (off4 > 2) v.z = kNegInf;
}
__device__ __forceinline__ void mask_tail(float4& v, int rem) {
  if (rem < 4) v.w = kNegInf;
  if (rem < 3) v.z = kNegInf;
  if (rem < 2) v.y = kNegInf;
}

template <int NS, int LPR, int CPL>
struct Rows {
  static constexpr int R = 32 / LPR;      // rows per warp pass
  static constexpr int NP = kTT / R;      // passes per tile
  static constexpr int Lpad = 32 * NS;
  static constexpr int NSL = Lpad / LPR;  // states per lane in the emission gather
  static constexpr int SEG = LPR * CPL;   // chunks per row segment
  static_assert(R <= kTT && NP * R == kTT, "tile must be a whole number of passes");

  const Problem& P;
  const FusedCfg& cfg;
  const Smem& S;
  const int lane, li, gi, wrow;
  const int64_t b;
  const int Tb, Lb, C;
  float* lse_arr;
  const uint64_t pol_stream;
  const float wgt;
  const uintptr_t base_addr;

  __device__ __forceinline__ Rows(const Problem& P_, const FusedCfg& cfg_, const Smem& S_, int lane_, int wrow_,
                                  int64_t b_, int Tb_, int Lb_, float wgt_)
      : P(P_), cfg(cfg_), S(S_), lane(lane_), li(lane_ & (LPR - 1)), gi(lane_ / LPR), wrow(wrow_), b(b_), Tb(Tb_),
        Lb(Lb_), C((int)P_.C), lse_arr(cfg_.lse_global ? cfg_.ws_lse + (size_t)b_ * P_.T : S_.lse),
        pol_stream(policy_evict_first()), wgt(wgt_), base_addr(reinterpret_cast<uintptr_t>(P_.logits)) {}

  // row t of this sequence; `slot_rows` = start of the ring slot holding the tile, i = row inside the tile
  __device__ __forceinline__ RowGeom geom(int t, const unsigned char* slot_rows, int i) const {
    RowGeom g;
    g.goff = ((int64_t)t * P.B + b) * C;
    g.off4 = (int)(((base_addr >> 2) + (uintptr_t)g.goff) & 3);
    g.nch = (g.off4 + C + 3) >> 2;
    g.rem = g.off4 + C - 4 * (g.nch - 1);
    g.srow = reinterpret_cast<const float4*>(slot_rows + (size_t)i * cfg.RS);
    return g;
  }

  __device__ __forceinline__ void load_seg(const RowGeom& g, bool act, int seg, float4 (&v)[CPL]) const {
    const float4* src = g.srow + seg * SEG + li;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int q = seg * SEG + li + c * LPR;
      v[c] = (act && q < g.nch) ? src[c * LPR] : make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
    }
  }
  // elements of the neighbouring rows (head of chunk 0 / tail of chunk nch-1) -> -inf
  template <bool kSingle>
  __device__ __forceinline__ void mask_seg(const RowGeom& g, int seg, bool first, bool last, float4 (&v)[CPL]) const {
    if (first && li == 0) mask_head(v[0], g.off4);
    if (last) {
      const int ql = g.nch - 1 - seg * SEG - li;  // tail chunk sits in slot c with c*LPR == ql
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        // with one segment the tail can only be in the last two slots (nch varies by <= 1 between rows)
        if (kSingle && c + 2 < CPL) continue;
        if (ql == c * LPR) mask_tail(v[c], g.rem);
      }
    }
  }

  __device__ __forceinline__ float group_max(float v) const {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  __device__ __forceinline__ float group_sum(float v) const {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }

  __device__ __forceinline__ void seg_max_sum(const float4 (&v)[CPL], float& m_run, float& s_run) const {
    float m = m_run;
#pragma unroll
    for (int c = 0; c < CPL; ++c) m = fmaxf(m, fmaxf(fmaxf(v[c].x, v[c].y), fmaxf(v[c].z, v[c].w)));
    if (m > kNegInf) {
      const float mb = m * kLog2e;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        s += ex2f(fmaf(v[c].x, kLog2e, -mb));
        s += ex2f(fmaf(v[c].y, kLog2e, -mb));
        s += ex2f(fmaf(v[c].z, kLog2e, -mb));
        s += ex2f(fmaf(v[c].w, kLog2e, -mb));
      }
      s_run = (m_run > kNegInf ? s_run * ex2f((m_run - m) * kLog2e) : 0.f) + s;
      m_run = m;
    }
  }

  // emissions p_t(s) = softmax(x_t)[label_s] for row i (gathered from the row's shared-memory copy) -> p-tile
  __device__ __forceinline__ void emit_row(const RowGeom& g, bool act, int i, float lse, float* ptile_buf) const {
    if (act) {
      const float* xr = reinterpret_cast<const float*>(g.srow) + g.off4;
      const float lb2 = lse * kLog2e;
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const int st = li + j * LPR;
        float pv = 0.f;
        if (st < Lb) pv = fmaxf(ex2f(fmaf(xr[S.lab[st]], kLog2e, -lb2)), kPMin);
        ptile_buf[i * Lpad + st] = pv;
      }
    }
  }

  // ---------------------------------------------------------------- phase 1: one tile
  __device__ __forceinline__ void forward_tile(int k, int nv, const unsigned char* slot_rows, float* ptile_buf) const {
#pragma unroll
    for (int pr = 0; pr < NP; ++pr) {
      const int i = pr * R + gi;
      const bool act = i < nv;
      const RowGeom g = geom(k * kTT + i, slot_rows, i);
      float m_run = kNegInf, s_run = 0.f;
      if (cfg.NSEG == 1) {
        float4 v[CPL];
        load_seg(g, act, 0, v);
        mask_seg<true>(g, 0, true, true, v);
        seg_max_sum(v, m_run, s_run);
      } else {
        for (int seg = 0; seg < cfg.NSEG; ++seg) {
          float4 v[CPL];
          load_seg(g, act, seg, v);
          mask_seg<false>(g, seg, seg == 0, seg == cfg.NSEG - 1, v);
          seg_max_sum(v, m_run, s_run);
        }
      }
      const float m = group_max(m_run);
      const float s = group_sum(m_run > kNegInf ? s_run * ex2f((m_run - m) * kLog2e) : 0.f);
      const float lse = m + logf(s);
      if (act && li == 0) lse_arr[k * kTT + i] = lse;
      emit_row(g, act, i, lse, ptile_buf);
    }
  }

  // ---------------------------------------------------------------- phase 2 stage A: emissions again
  __device__ __forceinline__ void emit_tile(int k, int nv, const unsigned char* slot_rows, float* ptile_buf) const {
#pragma unroll
    for (int pr = 0; pr < NP; ++pr) {
      const int i = pr * R + gi;
      const bool act = i < nv;
      const RowGeom g = geom(k * kTT + i, slot_rows, i);
      emit_row(g, act, i, act ? lse_arr[k * kTT + i] : 0.f, ptile_buf);
    }
  }

  // ---------------------------------------------------------------- gradient row pieces
  __device__ __forceinline__ void store_seg(float* grow, const RowGeom& g, int seg, const float4 (&v)[CPL]) const {
    float4* dst = reinterpret_cast<float4*>(grow - g.off4) + seg * SEG + li;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int q = seg * SEG + li + c * LPR;
      if (q < g.nch) {
        const bool head = (q == 0) && g.off4 != 0;
        const bool tail = (q == g.nch - 1) && g.rem != 4;
        if (!head && !tail) {
          stg_f4_hint(dst + c * LPR, v[c], pol_stream);
        } else {
          float* e = reinterpret_cast<float*>(dst + c * LPR);
          const int lo = head ? g.off4 : 0;
          const int hi = tail ? g.rem : 4;
          if (0 >= lo && 0 < hi) stg_f_hint(e + 0, v[c].x, pol_stream);
          if (1 >= lo && 1 < hi) stg_f_hint(e + 1, v[c].y, pol_stream);
          if (2 >= lo && 2 < hi) stg_f_hint(e + 2, v[c].z, pol_stream);
          if (3 >= lo && 3 < hi) stg_f_hint(e + 3, v[c].w, pol_stream);
        }
      }
    }
  }
  __device__ __forceinline__ void grad_seg(const float4* d4, const RowGeom& g, int seg, float lb2, float4 (&v)[CPL]) const {
    const float4* src = g.srow + seg * SEG + li;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int q = seg * SEG + li + c * LPR;
      if (q < g.nch) {
        const float4 x = src[c * LPR];
        const float4 d = d4[q];
        v[c].x = fmaf(ex2f(fmaf(x.x, kLog2e, -lb2)), wgt, -d.x);
        v[c].y = fmaf(ex2f(fmaf(x.y, kLog2e, -lb2)), wgt, -d.y);
        v[c].z = fmaf(ex2f(fmaf(x.z, kLog2e, -lb2)), wgt, -d.z);
        v[c].w = fmaf(ex2f(fmaf(x.w, kLog2e, -lb2)), wgt, -d.w);
      }
    }
  }
  __device__ __forceinline__ void scatter_add(float* dl, const RowGeom& g, bool act, int i, const float* gtile_buf) const {
    if (act) {
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const int st = li + j * LPR;
        if (st < Lb) atomicAdd(&dl[S.lab[st] + g.off4], gtile_buf[i * Lpad + st]);
      }
    }
  }
  __device__ __forceinline__ void scatter_clear(float* dl, const RowGeom& g, bool act) const {
    if (act) {
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const int st = li + j * LPR;
        if (st < Lb) dl[S.lab[st] + g.off4] = 0.f;
      }
    }
  }

  // ---------------------------------------------------------------- phase 2 stage B: one tile
  __device__ __forceinline__ void backward_tile(int k, int nv, const unsigned char* slot_rows, const float* gtile_buf) const {
    float* dl = S.delta + (wrow * R + gi) * cfg.Cd;
    const float4* d4 = reinterpret_cast<const float4*>(dl);
#pragma unroll
    for (int pr = 0; pr < NP; ++pr) {
      const int i = pr * R + gi;
      const bool act = i < nv;
      const RowGeom g = geom(k * kTT + i, slot_rows, i);
      scatter_add(dl, g, act, i, gtile_buf);
      __syncwarp();
      if (act) {
        const float lb2 = lse_arr[k * kTT + i] * kLog2e;
        float* grow = P.grad + g.goff;
        for (int seg = 0; seg < cfg.NSEG; ++seg) {
          float4 v[CPL];
          grad_seg(d4, g, seg, lb2, v);
          store_seg(grow, g, seg, v);
        }
      }
      __syncwarp();
      scatter_clear(dl, g, act);
      __syncwarp();
    }
  }

  // rows t in [t_begin, T) get an all-zero gradient (grads beyond input_length are exactly 0)
  __device__ __forceinline__ void zero_rows(int t_begin) const {
    float4 z[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) z[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t t = (int64_t)t_begin + wrow * R + gi; t < P.T; t += kNW * R) {
      const RowGeom g = geom((int)t, nullptr, 0);
      for (int seg = 0; seg < cfg.NSEG; ++seg) store_seg(P.grad + g.goff, g, seg, z);
    }
  }

  __device__ __forceinline__ void run() const {
    const int NT = (Tb + kTT - 1) / kTT;
    const int NBUF = cfg.NBUFP, NBUFG = cfg.NBUFG, NSLOT = cfg.NSLOT;
    const size_t slot_bytes = (size_t)kTT * cfg.RS;
    // ---- phase 1: tiles k = wrow, wrow+NW, ...
    for (int k = wrow; k < NT; k += kNW) {
      const int buf = k % NBUF;
      const int slot = k % NSLOT;
      if (k >= NBUF) mbar_wait(&S.pempty[buf], ((k / NBUF) - 1) & 1);
      mbar_wait(&S.sfull[slot], (k / NSLOT) & 1);
      forward_tile(k, min(kTT, Tb - k * kTT), S.ring + slot * slot_bytes, S.ptile + buf * (kTT * Lpad));
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&S.pfull[buf]);
        mbar_arrive(&S.sempty[slot]);
      }
    }
    if (P.grad == nullptr) return;
    if (Tb < P.T) zero_rows(Tb);
    // ---- phase 2: the same warp owns the same tiles (it wrote their lse values), walked downwards:
    // A(k) = emissions -> chain, B(k) = gradient rows once the chain has produced gamma(k).
    int k = NT - 1 - ((NT - 1 - wrow) % kNW + kNW) % kNW;  // largest k <= NT-1 with k % NW == wrow
    for (; k >= 0; k -= kNW) {
      const int j2 = NT - 1 - k;
      const int n = NT + j2;
      const int buf = n % NBUF, gbuf = j2 % NBUFG, slot = n % NSLOT;
      const int nv = min(kTT, Tb - k * kTT);
      const unsigned char* rows = S.ring + slot * slot_bytes;
      if (n >= NBUF) mbar_wait(&S.pempty[buf], ((n / NBUF) - 1) & 1);
      mbar_wait(&S.sfull[slot], (n / NSLOT) & 1);
      emit_tile(k, nv, rows, S.ptile + buf * (kTT * Lpad));
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.pfull[buf]);
      mbar_wait(&S.gfull[gbuf], (j2 / NBUFG) & 1);
      backward_tile(k, nv, rows, S.gtile + gbuf * (kTT * Lpad));
      if (lane == 0) {
        mbar_arrive(&S.gempty[gbuf]);
        mbar_arrive(&S.sempty[slot]);
      }
    }
  }
};

// ============================================================================ kernel
template <int NS, int LPR, int CPL>
__global__ void __launch_bounds__(kThreads, 3) nbctc_fused_kernel(const Problem P, const FusedCfg cfg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem S;
  S.pfull = reinterpret_cast<uint64_t*>(smem_raw + cfg.o_bar);
  S.pempty = S.pfull + kMaxBuf;
  S.gfull = S.pempty + kMaxBuf;
  S.gempty = S.gfull + kMaxBuf;
  S.sfull = S.gempty + kMaxBuf;
  S.sempty = S.sfull + kMaxSlot;
  S.lab = reinterpret_cast<int*>(smem_raw + cfg.o_lab);
  S.lse = reinterpret_cast<float*>(smem_raw + cfg.o_lse);
  S.ckpt = reinterpret_cast<double*>(smem_raw + cfg.o_ckpt);
  S.cke = reinterpret_cast<int*>(smem_raw + cfg.o_cke);
  S.ptile = reinterpret_cast<float*>(smem_raw + cfg.o_ptile);
  S.gtile = reinterpret_cast<float*>(smem_raw + cfg.o_gtile);
  S.atile = reinterpret_cast<double*>(smem_raw + cfg.o_atile);
  S.delta = reinterpret_cast<float*>(smem_raw + cfg.o_delta);
  S.ring = smem_raw + cfg.o_ring;

  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t Tb64 = P.in_len[b], Lb64 = P.tgt_len[b];
  bool ok = seq_feasible(Tb64, Lb64, P.T, P.Lmax);
  const int Tb = (int)Tb64, Lb = (int)Lb64;
  int bad = 0;
  if (ok) {
    for (int s = tid; s < 32 * NS; s += blockDim.x) {
      int l = 0;
      if (s < Lb) {
        l = P.labels[b * P.Lmax + s];
        if (l < 0 || l >= P.C) { bad = 1; l = 0; }
      }
      S.lab[s] = l;
    }
  }
  if (tid < kMaxBuf) {
    mbar_init(&S.pfull[tid], 1);
    mbar_init(&S.pempty[tid], 1);
    mbar_init(&S.gfull[tid], 1);
    mbar_init(&S.gempty[tid], 1);
  }
  if (tid < kMaxSlot) {
    mbar_init(&S.sfull[tid], 1);
    mbar_init(&S.sempty[tid], 1);
  }
  {
    const int nd = kNW * (32 / LPR) * cfg.Cd;
    for (int i = tid; i < nd; i += blockDim.x) S.delta[i] = 0.f;
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  bad = __syncthreads_or(bad);
  ok = ok && !bad;
  const float wgt = P.w_scalar * (P.seq_w ? P.seq_w[b] : 1.f);
  if (!ok) {
    if (tid == 0) P.loss[b] = INFINITY;
    if (P.grad != nullptr && warp >= 1 && warp <= kNW) {
      Rows<NS, LPR, CPL> rows(P, cfg, S, lane, warp - 1, b, 0, 0, wgt);
      rows.zero_rows(0);
    }
    return;
  }
  if (warp == 0) {
    chain_warp<NS>(P, cfg, S, lane, b, Tb, Lb, wgt);
  } else if (warp <= kNW) {
    Rows<NS, LPR, CPL> rows(P, cfg, S, lane, warp - 1, b, Tb, Lb, wgt);
    rows.run();
  } else {
    producer_warp(P, cfg, S, lane, b, Tb);
  }
}

template <int NS, int LPR, int CPL>
int launch_inst(const Problem& p, const FusedCfg& cfg, cudaStream_t stream) {
  auto kern = nbctc_fused_kernel<NS, LPR, CPL>;
  if (cfg.smem_bytes > 48 * 1024)
    NBCTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes));
  NBCTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  kern<<<(unsigned)p.B, kThreads, cfg.smem_bytes, stream>>>(p, cfg);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

template <int NS>
int launch_ns(const Problem& p, const FusedCfg& cfg, cudaStream_t stream) {
  if (cfg.LPR == 4) {
    switch (cfg.CPL) {
      case 1: return launch_inst<NS, 4, 1>(p, cfg, stream);
      case 2: return launch_inst<NS, 4, 2>(p, cfg, stream);
      case 3: return launch_inst<NS, 4, 3>(p, cfg, stream);
      case 4: return launch_inst<NS, 4, 4>(p, cfg, stream);
    }
  } else if (cfg.LPR == 8) {
    switch (cfg.CPL) {
      case 3: return launch_inst<NS, 8, 3>(p, cfg, stream);
      case 4: return launch_inst<NS, 8, 4>(p, cfg, stream);
      case 5: return launch_inst<NS, 8, 5>(p, cfg, stream);
      case 6: return launch_inst<NS, 8, 6>(p, cfg, stream);
      case 7: return launch_inst<NS, 8, 7>(p, cfg, stream);
      case 8: return launch_inst<NS, 8, 8>(p, cfg, stream);
    }
  } else if (cfg.LPR == 32) {
    switch (cfg.CPL) {
      case 3: return launch_inst<NS, 32, 3>(p, cfg, stream);
      case 4: return launch_inst<NS, 32, 4>(p, cfg, stream);
      case 6: return launch_inst<NS, 32, 6>(p, cfg, stream);
      case 8: return launch_inst<NS, 32, 8>(p, cfg, stream);
    }
  }
  set_error("no fused kernel instance for LPR=%d CPL=%d", cfg.LPR, cfg.CPL);
  return NBCTC_ERR_UNSUPPORTED;
}

}  // namespace fused
#endif  // __CUDACC__

}  // namespace nbctc
