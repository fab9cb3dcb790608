// Fused no-blank CTC forward+backward for sm_100a: ONE kernel reads the logits once from HBM,
// writes the gradient once, and keeps everything in between on chip.
//
// One CTA per sequence b, warp-specialised:
//   * "row" warps stream the (t,b) rows of the logits HBM -> registers with 128-bit loads on the
//     16-byte-aligned superset of each (only 4-byte aligned, C=157) row, several rows per warp
//     (LPR lanes per row).  Phase 1: row log-partition (NoBlankCTC.py:136) + per-state emission
//     gather (NoBlankCTC.py:96-102) -> p-tiles in shared memory.  Phase 2: the same rows again
//     (L2-resident, walked in reverse so the most recently read rows are re-read first) ->
//     w*(softmax - scatter(gamma)) -> 128-bit streaming stores.
//   * one "chain" warp runs the lattice recursions (NoBlankCTC.py:71-87) for the sequence, lane =
//     NS consecutive states, neighbour state through __shfl_up/down.  The state is kept in the
//     LINEAR domain in float64 with exact power-of-two rescaling once per tile, so a step is one
//     add and one multiply (no exp/log on the dependent chain) and sum_s alpha_t(s) beta_t(s) = Z
//     holds to 1e-13 -- gamma needs no per-row normalisation.  Phase 1 stores an alpha checkpoint
//     per tile of kTT steps; phase 2 re-plays alpha inside the tile next to the beta recursion
//     (NoBlankCTC.py:113-125 is the reference's commented-out backward pass; autograd does it there).
//   * tiles are handed over through shared memory with mbarriers (full/empty per buffer).
//
// Algorithmic HBM bytes per sequence: 2*4*T*C (+ labels); per real lattice cell 8*C/mean(L).
#include <cuda_runtime.h>

#include <algorithm>

#include "common.cuh"

namespace nbctc {
namespace {

constexpr int kTT = 8;        // time steps per tile
constexpr int kMaxBuf = 8;    // p-/gamma-tile ring depth upper bound
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kPMin = 7.52316385e-37f;  // 2^-120: emission floor (keeps 8 steps inside the f64 range)

struct FusedCfg {
  int NS;          // states per chain lane (1,2,4,8)
  int Lpad;        // 32*NS
  int LPR;         // lanes per row in the row warps (power of two)
  int NSEG;        // row segments of LPR*CPL chunks (1 unless the row is longer than 256 chunks)
  int NW;          // row warps per CTA
  int NBUF;        // tile ring depth
  int NTmax;       // ceil(T / kTT)
  int Cd;          // floats per scatter buffer
  int ckpt_global; // checkpoints live in the workspace instead of shared memory
  int lse_global;
  // shared-memory byte offsets
  uint32_t o_bar, o_lab, o_lse, o_ckpt, o_cke, o_ptile, o_gtile, o_atile, o_delta, smem_bytes;
  double* ws_ckpt;  // [B][NTmax][Lpad]
  int* ws_cke;      // [B][NTmax]
  float* ws_lse;    // [B][T]
};

// ---------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "NBCTC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra NBCTC_DONE;\n"
      "bra NBCTC_WAIT;\n"
      "NBCTC_DONE:\n"
      "}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ void stg_cs_f4(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void stg_cs_f(float* p, float v) { __stcs(p, v); }

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
  uint64_t* pfull;
  uint64_t* pempty;
  uint64_t* gfull;
  uint64_t* gempty;
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
__device__ __forceinline__ void alpha_step(double (&a)[NS], const double (&p)[NS], int lane) {
  double up = __shfl_up_sync(0xffffffffu, a[NS - 1], 1);
  if (lane == 0) up = 0.0;
#pragma unroll
  for (int j = NS - 1; j >= 1; --j) a[j] = (a[j] + a[j - 1]) * p[j];
  a[0] = (a[0] + up) * p[0];
}

template <int NS>
__device__ void chain_warp(const Problem& P, const FusedCfg& cfg, const Smem& S, int lane, int64_t b, int Tb, int Lb,
                           float wgt) {
  const int NT = (Tb + kTT - 1) / kTT;
  const int Lpad = cfg.Lpad, NBUF = cfg.NBUF;
  double* ck = cfg.ckpt_global ? cfg.ws_ckpt + ((size_t)b * cfg.NTmax) * Lpad : S.ckpt;
  int* cke = cfg.ckpt_global ? cfg.ws_cke + (size_t)b * cfg.NTmax : S.cke;
  double a[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) a[j] = 0.0;
  int Ea = 0;
  // ------------------------------------------------------------------ phase 1: alpha
  for (int k = 0; k < NT; ++k) {
    const int buf = k % NBUF;
    if (k > 0) {
      Ea += rescale_pow2<NS>(a);
#pragma unroll
      for (int j = 0; j < NS; ++j) ck[((size_t)k * NS + j) * 32 + lane] = a[j];
      if (lane == 0) cke[k] = Ea;
    }
    mbar_wait(&S.pfull[buf], (k / NBUF) & 1);
    const float* pt = S.ptile + (size_t)buf * kTT * Lpad + lane * NS;
    const int nv = min(kTT, Tb - k * kTT);
#pragma unroll
    for (int i = 0; i < kTT; ++i) {
      if (i < nv) {
        double p[NS];
        load_p<NS>(pt + i * Lpad, p);
        if (k == 0 && i == 0) {  // t = 0: only state 0 is reachable (NoBlankCTC.py:75,92-93)
#pragma unroll
          for (int j = 0; j < NS; ++j) a[j] = (lane == 0 && j == 0) ? p[0] : 0.0;
        } else {
          alpha_step<NS>(a, p, lane);
        }
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
  if (lane == 0) {
    // -log Z; Z = zhat * 2^Ez
    float l = (zhat > 0.0) ? (float)(-(log(zhat) + (double)Ez * 0.6931471805599453)) : INFINITY;
    P.loss[b] = l;
  }
  if (P.grad == nullptr) return;
  const double zinv = (zhat > 0.0) ? (double)wgt / zhat : 0.0;  // w folded into gamma
  // ------------------------------------------------------------------ phase 2: beta, gamma
  double u[NS];  // beta_{t+1}(s) * p_{t+1}(s), scaled by 2^-Eb
#pragma unroll
  for (int j = 0; j < NS; ++j) u[j] = 0.0;
  int Eb = 0;
  for (int j2 = 0; j2 < NT; ++j2) {
    const int k = NT - 1 - j2;
    const int n = NT + j2;
    const int buf = n % NBUF;
    const int gbuf = j2 % NBUF;
    int EaK = 0;
    if (k == 0) {
#pragma unroll
      for (int j = 0; j < NS; ++j) a[j] = 0.0;
    } else {
#pragma unroll
      for (int j = 0; j < NS; ++j) a[j] = ck[((size_t)k * NS + j) * 32 + lane];
      EaK = cke[k];
    }
    mbar_wait(&S.pfull[buf], (n / NBUF) & 1);
    const float* pt = S.ptile + (size_t)buf * kTT * Lpad + lane * NS;
    const int nv = min(kTT, Tb - k * kTT);
    // replay alpha inside the tile
#pragma unroll
    for (int i = 0; i < kTT; ++i) {
      if (i < nv) {
        double p[NS];
        load_p<NS>(pt + i * Lpad, p);
        if (k == 0 && i == 0) {
#pragma unroll
          for (int j = 0; j < NS; ++j) a[j] = (lane == 0 && j == 0) ? p[0] : 0.0;
        } else {
          alpha_step<NS>(a, p, lane);
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) S.atile[((size_t)i * NS + j) * 32 + lane] = a[j];
      }
    }
    // gamma = alpha * beta * w / Z, exponents split over two exact power-of-two factors
    const int d = EaK + Eb - Ez;
    const double s1 = pow2i(d / 2);
    const double s2 = pow2i(d - d / 2) * zinv;
    if (j2 >= NBUF) mbar_wait(&S.gempty[gbuf], ((j2 / NBUF) - 1) & 1);
    float* gt = S.gtile + (size_t)gbuf * kTT * Lpad + lane * NS;
#pragma unroll
    for (int ii = 0; ii < kTT; ++ii) {
      const int i = kTT - 1 - ii;
      if (i < nv) {
        const int t = k * kTT + i;
        double p[NS], beta[NS];
        load_p<NS>(pt + i * Lpad, p);
        if (t == Tb - 1) {
#pragma unroll
          for (int j = 0; j < NS; ++j) beta[j] = (lane * NS + j == sl) ? 1.0 : 0.0;
        } else {
          double dn = __shfl_down_sync(0xffffffffu, u[0], 1);
          if (lane == 31) dn = 0.0;
#pragma unroll
          for (int j = 0; j < NS - 1; ++j) beta[j] = u[j] + u[j + 1];
          beta[NS - 1] = u[NS - 1] + dn;
        }
        float g[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          const double al = S.atile[((size_t)i * NS + j) * 32 + lane];
          g[j] = (float)((al * s1) * (beta[j] * s2));
          u[j] = beta[j] * p[j];
        }
        if constexpr (NS == 1) {
          gt[i * Lpad] = g[0];
        } else if constexpr (NS == 2) {
          *reinterpret_cast<float2*>(gt + i * Lpad) = make_float2(g[0], g[1]);
        } else {
#pragma unroll
          for (int j = 0; j < NS; j += 4)
            *reinterpret_cast<float4*>(gt + i * Lpad + j) = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
        }
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

// ============================================================================ row warps
struct RowGeom {
  const float4* base;  // 16-byte aligned start of the chunked row
  int off4;            // floats between base and the row's first element (0..3)
  int nch;             // 16-byte chunks covering the row
};
__device__ __forceinline__ RowGeom row_geom(const float* row_ptr, int C) {
  RowGeom g;
  g.off4 = (int)((reinterpret_cast<uintptr_t>(row_ptr) >> 2) & 3);
  g.base = reinterpret_cast<const float4*>(row_ptr - g.off4);
  g.nch = (g.off4 + C + 3) >> 2;
  return g;
}

template <int CPL>
__device__ __forceinline__ void load_row(const RowGeom& g, bool active, int q0, int LPR, float4 (&v)[CPL]) {
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int q = q0 + i * LPR;
    if (active && q < g.nch)
      v[i] = ldg_f4(g.base + q);
    else
      v[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  }
}

// elements that belong to the neighbouring rows (head of chunk 0, tail of the last chunk) -> -inf
template <int CPL>
__device__ __forceinline__ void mask_row(const RowGeom& g, int C, int q0, int LPR, float4 (&v)[CPL]) {
  if (q0 == 0 && g.off4 != 0) {
    v[0].x = -INFINITY;
    if (g.off4 > 1) v[0].y = -INFINITY;
    if (g.off4 > 2) v[0].z = -INFINITY;
  }
  const int rem = (g.off4 + C) & 3;  // valid elements in the last chunk (0 = all four)
  if (rem != 0) {
    const int ql = g.nch - 1;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      if (q0 + i * LPR == ql) {
        v[i].w = -INFINITY;
        if (rem < 3) v[i].z = -INFINITY;
        if (rem < 2) v[i].y = -INFINITY;
      }
    }
  }
}

__device__ __forceinline__ float group_max(float v, int LPR) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    if (o < LPR) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group_sum(float v, int LPR) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    if (o < LPR) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// write one (possibly partial) row of gradient chunks
template <int CPL>
__device__ __forceinline__ void store_row(float* grow, const RowGeom& g, int C, int q0, int LPR, const float4 (&v)[CPL]) {
  float4* gb = reinterpret_cast<float4*>(grow - g.off4);
  const int rem = (g.off4 + C) & 3;
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int q = q0 + i * LPR;
    if (q < g.nch) {
      const bool head = (q == 0 && g.off4 != 0);
      const bool tail = (q == g.nch - 1 && rem != 0);
      if (!head && !tail) {
        stg_cs_f4(gb + q, v[i]);
      } else {
        float* e = reinterpret_cast<float*>(gb + q);
        const int lo = head ? g.off4 : 0;
        const int hi = tail ? rem : 4;
        if (0 >= lo && 0 < hi) stg_cs_f(e + 0, v[i].x);
        if (1 >= lo && 1 < hi) stg_cs_f(e + 1, v[i].y);
        if (2 >= lo && 2 < hi) stg_cs_f(e + 2, v[i].z);
        if (3 >= lo && 3 < hi) stg_cs_f(e + 3, v[i].w);
      }
    }
  }
}

// Phase-1 work of one row warp on tile k: lse + emissions for up to kTT rows.
template <int CPL>
__device__ __forceinline__ void rows_forward_tile(const Problem& P, const FusedCfg& cfg, const Smem& S, float* lse_arr,
                                                  int lane, int64_t b, int k, int nv, int Lb, float* ptile_buf) {
  const int LPR = cfg.LPR, R = 32 / LPR, li = lane & (LPR - 1), gi = lane / LPR;
  const int C = (int)P.C;
  const int Lpad = cfg.Lpad;
  for (int r0 = 0; r0 < nv; r0 += R) {
    const int i = r0 + gi;  // row inside the tile
    const bool active = i < nv;
    const int64_t t = (int64_t)k * kTT + i;
    const float* xrow = P.logits + (t * P.B + b) * C;
    RowGeom g = row_geom(xrow, C);
    // lane-local running (max, sum) over the row segments, merged across the group at the end
    float m_run = -INFINITY, s_run = 0.f;
    for (int seg = 0; seg < cfg.NSEG; ++seg) {
      const int q0 = seg * LPR * CPL + li;
      float4 v[CPL];
      load_row<CPL>(g, active, q0, LPR, v);
      mask_row<CPL>(g, C, q0, LPR, v);
      float m = m_run;
#pragma unroll
      for (int c = 0; c < CPL; ++c) m = fmaxf(m, fmaxf(fmaxf(v[c].x, v[c].y), fmaxf(v[c].z, v[c].w)));
      if (m > -INFINITY) {
        const float mb = m * kLog2e;
        float sseg = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          sseg += ex2f(fmaf(v[c].x, kLog2e, -mb));
          sseg += ex2f(fmaf(v[c].y, kLog2e, -mb));
          sseg += ex2f(fmaf(v[c].z, kLog2e, -mb));
          sseg += ex2f(fmaf(v[c].w, kLog2e, -mb));
        }
        s_run = (m_run > -INFINITY ? s_run * ex2f((m_run - m) * kLog2e) : 0.f) + sseg;
        m_run = m;
      }
    }
    const float m = group_max(m_run, LPR);
    float s = (m_run > -INFINITY) ? s_run * ex2f((m_run - m) * kLog2e) : 0.f;
    s = group_sum(s, LPR);
    const float lse = m + logf(s);
    if (active) {
      if (li == 0) lse_arr[t] = lse;
      const float lb2 = lse * kLog2e;
      for (int st = li; st < Lpad; st += LPR) {
        float pv = 0.f;
        if (st < Lb) pv = fmaxf(ex2f(fmaf(__ldg(xrow + S.lab[st]), kLog2e, -lb2)), kPMin);
        ptile_buf[i * Lpad + st] = pv;
      }
    }
  }
}

// Phase-2 stage A: emissions of tile k again (row constants are known now).
__device__ __forceinline__ void rows_emit_tile(const Problem& P, const FusedCfg& cfg, const Smem& S, const float* lse_arr,
                                               int lane, int64_t b, int k, int nv, int Lb, float* ptile_buf) {
  const int C = (int)P.C;
  const int Lpad = cfg.Lpad;
  // all 32 lanes walk the (row, state) pairs of the tile
  const int total = nv * Lpad;
  for (int idx = lane; idx < total; idx += 32) {
    const int i = idx / Lpad, st = idx - i * Lpad;
    const int64_t t = (int64_t)k * kTT + i;
    float pv = 0.f;
    if (st < Lb) {
      const float* xrow = P.logits + (t * P.B + b) * C;
      pv = fmaxf(ex2f(fmaf(__ldg(xrow + S.lab[st]), kLog2e, -lse_arr[t] * kLog2e)), kPMin);
    }
    ptile_buf[idx] = pv;
  }
}

// Phase-2 stage B: gradient rows of tile k.
template <int CPL>
__device__ __forceinline__ void rows_backward_tile(const Problem& P, const FusedCfg& cfg, const Smem& S,
                                                   const float* lse_arr, int lane, int wrow, int64_t b, int k, int nv,
                                                   int Lb, const float* gtile_buf, float wgt) {
  const int LPR = cfg.LPR, R = 32 / LPR, li = lane & (LPR - 1), gi = lane / LPR;
  const int C = (int)P.C;
  const int Lpad = cfg.Lpad;
  float* dl = S.delta + ((size_t)wrow * R + gi) * cfg.Cd;
  for (int r0 = 0; r0 < nv; r0 += R) {
    const int i = r0 + gi;
    const bool active = i < nv;
    const int64_t t = (int64_t)k * kTT + i;
    const int64_t roff = (t * P.B + b) * C;
    const float* xrow = P.logits + roff;
    RowGeom g = row_geom(xrow, C);
    if (active) {
      for (int st = li; st < Lb; st += LPR) atomicAdd(&dl[S.lab[st] + g.off4], gtile_buf[i * Lpad + st]);
    }
    __syncwarp();
    if (active) {
      const float lb2 = lse_arr[t] * kLog2e;
      const float4* d4 = reinterpret_cast<const float4*>(dl);
      for (int seg = 0; seg < cfg.NSEG; ++seg) {
        const int q0 = seg * LPR * CPL + li;
        float4 v[CPL];
        load_row<CPL>(g, true, q0, LPR, v);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const int q = q0 + c * LPR;
          if (q < g.nch) {
            const float4 d = d4[q];
            v[c].x = fmaf(ex2f(fmaf(v[c].x, kLog2e, -lb2)), wgt, -d.x);
            v[c].y = fmaf(ex2f(fmaf(v[c].y, kLog2e, -lb2)), wgt, -d.y);
            v[c].z = fmaf(ex2f(fmaf(v[c].z, kLog2e, -lb2)), wgt, -d.z);
            v[c].w = fmaf(ex2f(fmaf(v[c].w, kLog2e, -lb2)), wgt, -d.w);
          }
        }
        store_row<CPL>(P.grad + roff, g, C, q0, LPR, v);
      }
    }
    __syncwarp();
    if (active) {
      for (int st = li; st < Lb; st += LPR) dl[S.lab[st] + g.off4] = 0.f;
    }
  }
}

// rows t in [t_begin, T) of sequence b get an all-zero gradient (NoBlankCTC quirk: grads beyond input_length are 0)
template <int CPL>
__device__ __forceinline__ void rows_zero(const Problem& P, const FusedCfg& cfg, int lane, int wrow, int64_t b,
                                          int64_t t_begin) {
  const int LPR = cfg.LPR, R = 32 / LPR, li = lane & (LPR - 1), gi = lane / LPR;
  const int C = (int)P.C;
  float4 z[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) z[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t t = t_begin + (int64_t)wrow * R + gi; t < P.T; t += (int64_t)cfg.NW * R) {
    const int64_t roff = (t * P.B + b) * C;
    RowGeom g = row_geom(P.logits + roff, C);
    for (int seg = 0; seg < cfg.NSEG; ++seg) store_row<CPL>(P.grad + roff, g, C, seg * LPR * CPL + li, LPR, z);
  }
}

template <int CPL>
__device__ void row_warp(const Problem& P, const FusedCfg& cfg, const Smem& S, int lane, int wrow, int64_t b, int Tb,
                         int Lb, float wgt) {
  const int NT = (Tb + kTT - 1) / kTT;
  const int NBUF = cfg.NBUF, NW = cfg.NW, Lpad = cfg.Lpad;
  float* lse_arr = cfg.lse_global ? cfg.ws_lse + (size_t)b * P.T : S.lse;
  // ---- phase 1: tiles k = wrow, wrow+NW, ...
  for (int k = wrow; k < NT; k += NW) {
    const int buf = k % NBUF;
    if (k >= NBUF) mbar_wait(&S.pempty[buf], ((k / NBUF) - 1) & 1);
    const int nv = min(kTT, Tb - k * kTT);
    rows_forward_tile<CPL>(P, cfg, S, lse_arr, lane, b, k, nv, Lb, S.ptile + (size_t)buf * kTT * Lpad);
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.pfull[buf]);
  }
  if (P.grad == nullptr) return;
  if (Tb < P.T) rows_zero<CPL>(P, cfg, lane, wrow, b, Tb);
  // ---- phase 2: the same warp owns the same tiles (its own lse values), walked downwards.
  // Order per warp: A(k0), A(k0-NW), B(k0), A(k0-2NW), B(k0-NW), ... so the chain always has a tile ahead.
  int kA = -1;
  for (int k = NT - 1; k >= 0; --k)
    if (k % NW == wrow) { kA = k; break; }
  int kB = kA;
  auto stage_a = [&](int k) {
    const int n = NT + (NT - 1 - k);
    const int buf = n % NBUF;
    if (n >= NBUF) mbar_wait(&S.pempty[buf], ((n / NBUF) - 1) & 1);
    const int nv = min(kTT, Tb - k * kTT);
    rows_emit_tile(P, cfg, S, lse_arr, lane, b, k, nv, Lb, S.ptile + (size_t)buf * kTT * Lpad);
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.pfull[buf]);
  };
  auto stage_b = [&](int k) {
    const int j2 = NT - 1 - k;
    const int gbuf = j2 % NBUF;
    mbar_wait(&S.gfull[gbuf], (j2 / NBUF) & 1);
    const int nv = min(kTT, Tb - k * kTT);
    rows_backward_tile<CPL>(P, cfg, S, lse_arr, lane, wrow, b, k, nv, Lb, S.gtile + (size_t)gbuf * kTT * Lpad, wgt);
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.gempty[gbuf]);
  };
  if (kA >= 0) {
    stage_a(kA);
    kA -= NW;
  }
  while (kB >= 0) {
    if (kA >= 0) {
      stage_a(kA);
      kA -= NW;
    }
    stage_b(kB);
    kB -= NW;
  }
}

// ============================================================================ kernel
template <int NS, int CPL>
__global__ void __launch_bounds__(32 * 9) nbctc_fused_kernel(Problem P, FusedCfg cfg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem S;
  S.pfull = reinterpret_cast<uint64_t*>(smem_raw + cfg.o_bar);
  S.pempty = S.pfull + kMaxBuf;
  S.gfull = S.pempty + kMaxBuf;
  S.gempty = S.gfull + kMaxBuf;
  S.lab = reinterpret_cast<int*>(smem_raw + cfg.o_lab);
  S.lse = reinterpret_cast<float*>(smem_raw + cfg.o_lse);
  S.ckpt = reinterpret_cast<double*>(smem_raw + cfg.o_ckpt);
  S.cke = reinterpret_cast<int*>(smem_raw + cfg.o_cke);
  S.ptile = reinterpret_cast<float*>(smem_raw + cfg.o_ptile);
  S.gtile = reinterpret_cast<float*>(smem_raw + cfg.o_gtile);
  S.atile = reinterpret_cast<double*>(smem_raw + cfg.o_atile);
  S.delta = reinterpret_cast<float*>(smem_raw + cfg.o_delta);

  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t Tb64 = P.in_len[b], Lb64 = P.tgt_len[b];
  bool ok = seq_feasible(Tb64, Lb64, P.T, P.Lmax);
  const int Tb = (int)Tb64, Lb = (int)Lb64;
  int bad = 0;
  if (ok) {
    for (int s = tid; s < cfg.Lpad; s += blockDim.x) {
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
  {
    const int nd = cfg.NW * (32 / cfg.LPR) * cfg.Cd;
    for (int i = tid; i < nd; i += blockDim.x) S.delta[i] = 0.f;
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  bad = __syncthreads_or(bad);
  ok = ok && !bad;
  float wgt = P.w_scalar * (P.seq_w ? P.seq_w[b] : 1.f);
  if (!ok) {
    if (tid == 0) P.loss[b] = INFINITY;
    if (P.grad != nullptr && warp > 0) rows_zero<CPL>(P, cfg, lane, warp - 1, b, 0);
    return;
  }
  if (warp == 0)
    chain_warp<NS>(P, cfg, S, lane, b, Tb, Lb, wgt);
  else
    row_warp<CPL>(P, cfg, S, lane, warp - 1, b, Tb, Lb, wgt);
}

// ---------------------------------------------------------------------------- host side
struct Plan {
  bool ok;
  FusedCfg cfg;
  int CPL;
};

Plan make_plan(int64_t T, int64_t B, int64_t C, int64_t Lmax, const void* logits_ptr) {
  Plan pl{};
  pl.ok = false;
  if (Lmax > 256 || T > (int64_t)1 << 30 || C > (int64_t)1 << 24) return pl;
  FusedCfg& c = pl.cfg;
  c.NS = Lmax <= 32 ? 1 : Lmax <= 64 ? 2 : Lmax <= 128 ? 4 : 8;
  c.Lpad = 32 * c.NS;
  // chunks per row: rows of a 16B-aligned tensor with C % 4 == 0 are themselves aligned
  const int off_base = (int)((reinterpret_cast<uintptr_t>(logits_ptr) >> 2) & 3);
  const int64_t nch = (C % 4 == 0) ? (off_base + C + 3) / 4 : (3 + C + 3) / 4;
  // lanes per row / chunks per lane: least padding, then more lanes; rows longer than 256 chunks
  // are walked in NSEG segments of 32 lanes x 8 chunks
  c.NSEG = 1;
  int best_lpr = 32, best_cpl = 8, best_waste = 1 << 30;
  if (nch > 256) {
    c.NSEG = (int)((nch + 255) / 256);
    best_waste = 0;
  }
  for (int lpr = 4; lpr <= 32 && nch <= 256; lpr *= 2) {
    const int cpl = (int)((nch + lpr - 1) / lpr);
    if (cpl > 8) continue;
    const int waste = lpr * cpl - (int)nch;
    if (waste < best_waste || (waste == best_waste && lpr > best_lpr)) {
      best_waste = waste; best_lpr = lpr; best_cpl = cpl;
    }
    if (cpl <= 4) break;  // do not spread a short row over more lanes than needed
  }
  c.LPR = best_lpr;
  pl.CPL = best_cpl;
  c.NW = 4;
  c.NTmax = (int)((T + kTT - 1) / kTT);
  c.Cd = (int)((C + 3 + 3) / 4 * 4 + 4);
  // shared-memory carve-up
  const size_t budget = 200 * 1024;
  auto layout = [&](int nbuf, bool ck_glob, bool lse_glob) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 16); return (uint32_t)o; };
    c.NBUF = nbuf; c.ckpt_global = ck_glob; c.lse_global = lse_glob;
    c.o_bar = take(sizeof(uint64_t) * 4 * kMaxBuf);
    c.o_lab = take(sizeof(int) * c.Lpad);
    c.o_lse = take(lse_glob ? 16 : sizeof(float) * T);
    c.o_ckpt = take(ck_glob ? 16 : sizeof(double) * (size_t)c.NTmax * c.Lpad);
    c.o_cke = take(ck_glob ? 16 : sizeof(int) * (size_t)c.NTmax);
    c.o_ptile = take(sizeof(float) * (size_t)nbuf * kTT * c.Lpad);
    c.o_gtile = take(sizeof(float) * (size_t)nbuf * kTT * c.Lpad);
    c.o_atile = take(sizeof(double) * (size_t)kTT * c.Lpad);
    c.o_delta = take(sizeof(float) * (size_t)c.NW * (32 / c.LPR) * c.Cd);
    c.smem_bytes = (uint32_t)off;
    return off;
  };
  // prefer everything in shared memory; at most ~45 KB per CTA keeps 4-5 CTAs per SM
  const int nbuf_pref = 2 * c.NW;
  bool placed = false;
  for (int pass = 0; pass < 4 && !placed; ++pass) {
    const bool ckg = pass >= 1, lsg = pass >= 2;
    for (int nbuf = nbuf_pref; nbuf >= c.NW + 1; --nbuf) {
      size_t need = layout(nbuf, ckg, lsg);
      const size_t cap = (pass == 0) ? 56 * 1024 : budget;
      if (need <= cap) { placed = true; break; }
      if (pass == 0) break;
    }
  }
  if (!placed) return pl;
  pl.ok = true;
  return pl;
}

size_t plan_ws_bytes(const Plan& pl, int64_t T, int64_t B) {
  size_t off = 256;
  if (pl.cfg.ckpt_global) {
    off = align_up(off + sizeof(double) * (size_t)B * pl.cfg.NTmax * pl.cfg.Lpad, 256);
    off = align_up(off + sizeof(int) * (size_t)B * pl.cfg.NTmax, 256);
  }
  if (pl.cfg.lse_global) off = align_up(off + sizeof(float) * (size_t)B * T, 256);
  return off;
}

template <int NS, int CPL>
int launch_inst(const Problem& p, const FusedCfg& cfg, cudaStream_t stream) {
  auto kern = nbctc_fused_kernel<NS, CPL>;
  if (cfg.smem_bytes > 48 * 1024)
    NBCTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes));
  kern<<<(unsigned)p.B, 32 * (1 + cfg.NW), cfg.smem_bytes, stream>>>(p, cfg);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

template <int NS>
int launch_ns(const Problem& p, const FusedCfg& cfg, int CPL, cudaStream_t stream) {
  switch (CPL) {
    case 1: return launch_inst<NS, 1>(p, cfg, stream);
    case 2: return launch_inst<NS, 2>(p, cfg, stream);
    case 3: return launch_inst<NS, 3>(p, cfg, stream);
    case 4: return launch_inst<NS, 4>(p, cfg, stream);
    case 5: return launch_inst<NS, 5>(p, cfg, stream);
    case 6: return launch_inst<NS, 6>(p, cfg, stream);
    case 7: return launch_inst<NS, 7>(p, cfg, stream);
    default: return launch_inst<NS, 8>(p, cfg, stream);
  }
}

}  // namespace

bool fused_supported(int64_t T, int64_t B, int64_t C, int64_t Lmax, bool binary) {
  if (binary) return false;
  // alignment-independent answer: plan with the worst-case base alignment
  return make_plan(T, B, C, Lmax, reinterpret_cast<const void*>(uintptr_t(C % 4 == 0 ? 0 : 4))).ok &&
         make_plan(T, B, C, Lmax, reinterpret_cast<const void*>(uintptr_t(12))).ok;
}

size_t fused_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax, bool binary) {
  (void)binary;
  Plan a = make_plan(T, B, C, Lmax, reinterpret_cast<const void*>(uintptr_t(0)));
  Plan b = make_plan(T, B, C, Lmax, reinterpret_cast<const void*>(uintptr_t(12)));
  return std::max(plan_ws_bytes(a, T, B), plan_ws_bytes(b, T, B));
}

int fused_launch(const Problem& p, bool binary, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (binary) {
    set_error("fused binary path not available");
    return NBCTC_ERR_UNSUPPORTED;
  }
  Plan pl = make_plan(p.T, p.B, p.C, p.Lmax, p.logits);
  if (!pl.ok) {
    set_error("shape not supported by the fused kernel");
    return NBCTC_ERR_UNSUPPORTED;
  }
  const size_t need = plan_ws_bytes(pl, p.T, p.B);
  if ((pl.cfg.ckpt_global || pl.cfg.lse_global) && (ws == nullptr || ws_bytes < need)) {
    set_error("workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    return NBCTC_ERR_WORKSPACE;
  }
  char* w = static_cast<char*>(ws);
  size_t off = 256;
  if (pl.cfg.ckpt_global) {
    pl.cfg.ws_ckpt = reinterpret_cast<double*>(w + off);
    off = align_up(off + sizeof(double) * (size_t)p.B * pl.cfg.NTmax * pl.cfg.Lpad, 256);
    pl.cfg.ws_cke = reinterpret_cast<int*>(w + off);
    off = align_up(off + sizeof(int) * (size_t)p.B * pl.cfg.NTmax, 256);
  }
  if (pl.cfg.lse_global) pl.cfg.ws_lse = reinterpret_cast<float*>(w + off);
  switch (pl.cfg.NS) {
    case 1: return launch_ns<1>(p, pl.cfg, pl.CPL, stream);
    case 2: return launch_ns<2>(p, pl.cfg, pl.CPL, stream);
    case 4: return launch_ns<4>(p, pl.cfg, pl.CPL, stream);
    default: return launch_ns<8>(p, pl.cfg, pl.CPL, stream);
  }
}

}  // namespace nbctc
