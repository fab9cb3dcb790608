// Fused no-blank CTC forward+backward kernel for sm_100a (overview in nbctc_fused.cu).
//
// One CTA per sequence b, warp-specialised, all hand-offs through shared memory + mbarriers:
//   warp 0        chain warp.  16 lanes x NS states hold the lattice state in float64 (linear domain, exact
//                 power-of-two rescaling per tile).  Lanes 0-15 run alpha; in phase 2 lanes 16-31 run the beta
//                 recursion IN THE SAME INSTRUCTIONS (beta is stored in reversed state order so both halves shift
//                 the same way), while lanes 0-15 replay alpha inside the tile from the phase-1 checkpoint.
//   warps 1..NW   row warps.  LPR lanes per (t,b) row, CPL 16-byte chunks per lane: row log-partition + emission
//                 gather (phase 1), softmax - scatter(gamma) and 128-bit streaming stores (phase 2).
//   warp NW+1     TMA producer.  One lane issues cp.async.bulk copies of the rows (16-byte aligned superset of the
//                 4-byte aligned rows) into a shared-memory ring: phase 1 upwards with an L2 evict_last policy,
//                 phase 2 downwards (most recently read rows first -> L2 hits) with evict_first.
// Template parameters: NS (chain states per lane, Lmax <= 16*NS), LPR, CPL (row geometry).
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
  int NS, Lpad;      // chain: NS states per lane, 16 lanes per direction, Lpad = 16*NS
  int LPR, CPL, NSEG;
  int NBUFP, NBUFG;  // p-tile / gamma-tile ring depths
  int NSLOT;         // row-tile ring depth (TMA bulk copies land here)
  int RS;            // bytes per row slot = 16 * max chunks per row
  int NTmax;         // ceil(T / kTT)
  int ckpt_global;   // checkpoints live in the workspace instead of shared memory
  int lse_global;
  uint32_t o_bar, o_lab, o_lse, o_ckpt, o_cke, o_ptile, o_gtile, o_abtile, o_ring, smem_bytes;
  double* ws_ckpt;   // [B][NTmax][Lpad]
  int* ws_cke;       // [B][NTmax]
  float* ws_lse;     // [B][T]
  const float* logits_end;  // one past the last logit (bulk copies never read past its 16-byte round-up)
};

int launch_fused_ns2(const Problem& p, const FusedCfg& cfg, cudaStream_t stream);
int launch_fused_ns4(const Problem& p, const FusedCfg& cfg, cudaStream_t stream);
int launch_fused_ns8(const Problem& p, const FusedCfg& cfg, cudaStream_t stream);
int launch_fused_ns16(const Problem& p, const FusedCfg& cfg, cudaStream_t stream);

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
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
// Wait for the phase with the given parity.  A protocol bug would otherwise hang the GPU: after ~2 s of polling
// the kernel traps instead (the host sees a launch failure), which is never reached in a correct run.
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t bar = smem_u32(b);
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try_wait(bar, parity)) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 2000000000ull) {
      printf("nbctc: mbarrier wait timed out (block %d warp %d bar+%u parity %u)\n", (int)blockIdx.x, (int)(threadIdx.x >> 5),
             bar, parity);
      __trap();
    }
  }
}
// Monotonic shared-memory counters guard the rings that have several waiting warps: an mbarrier parity wait is only
// sound while the waiter is at most one phase ahead of the barrier, which a warp that skips ahead in the tile order
// (phase-1 -> phase-2 hand-over) cannot guarantee.  Publisher: fence + store; waiter: spin + fence.
__device__ __forceinline__ void count_publish(volatile int* c, int v) {
  __threadfence_block();
  *c = v;
}
__device__ __forceinline__ void count_wait(volatile int* c, int target) {
  if (*c >= target) { __threadfence_block(); return; }
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (*c < target) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 2000000000ull) {
      printf("nbctc: counter wait timed out (block %d warp %d target %d have %d)\n", (int)blockIdx.x,
             (int)(threadIdx.x >> 5), target, *c);
      __trap();
    }
  }
  __threadfence_block();
}
// TMA bulk copy global -> shared (1-D, 16-byte aligned, size multiple of 16), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, uint64_t src_gmem, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          dst_smem),
      "l"(src_gmem), "r"(bytes), "r"(bar), "l"(pol)
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
__device__ __forceinline__ void stg_f4_hint(float4* ptr, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void stg_f_hint(float* ptr, float v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(ptr), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ double pow2i(int e) {  // exact 2^e, e clamped to the normal range
  e = max(-1022, min(1023, e));
  return __hiloint2double((1023 + e) << 20, 0);
}

// Optional role profiler (compile with -DNBCTC_PROF): per-warp cycle counters split into wait / work buckets,
// dumped to a debug buffer.  Compiled out of the product build.
#ifdef NBCTC_PROF
extern __device__ long long* g_nbctc_prof;
#define PROF_DECL long long prof_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long prof_t0_ = clock64();
#define PROF_SCOPE(i, stmt) { const long long t_ = clock64(); stmt; prof_[i] += clock64() - t_; }
#define PROF_DUMP(warp_)                                                                          \
  if (lane == 0 && g_nbctc_prof != nullptr) {                                                     \
    prof_[7] = clock64() - prof_t0_;                                                              \
    for (int i_ = 0; i_ < 8; ++i_) g_nbctc_prof[((size_t)b * 6 + (warp_)) * 8 + i_] = prof_[i_];  \
  }
#else
#define PROF_DECL
#define PROF_SCOPE(i, stmt) { stmt; }
#define PROF_DUMP(warp_)
#endif

struct Smem {
  uint64_t *pfull, *gfull, *gempty, *sfull, *sempty;
  volatile int* cnt;  // [0] p-tiles consumed by the chain, [1] row tiles issued by the producer (monotonic)
  int* lab;
  float* lse;
  double* ckpt;
  int* cke;
  float* ptile;
  float* gtile;
  double* abtile;
  unsigned char* ring;
};

// ============================================================================ chain warp
// rescale the NS states of each 16-lane half by the exact power of two of the half's largest value
template <int NS>
__device__ __forceinline__ int rescale_half(double (&v)[NS]) {
  double m = v[0];
#pragma unroll
  for (int j = 1; j < NS; ++j) m = fmax(m, v[j]);
  int hi = __double2hiint(m);  // values are >= 0, so the high word orders like the value
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  const int ex = hi >> 20;
  if (ex == 0 || ex >= 0x7ff) return 0;
  const int e = ex - 1023;
  const double sc = __hiloint2double((1023 - e) << 20, 0);  // exact 2^-e
#pragma unroll
  for (int j = 0; j < NS; ++j) v[j] *= sc;
  return e;
}

// emissions of the lane's NS states for one row of a p-tile; the beta half reads them in reversed state order
template <int NS>
__device__ __forceinline__ void load_p(const float* row, int hl, bool rev, double (&p)[NS]) {
  const float* src = row + (rev ? (16 - 1 - hl) * NS : hl * NS);
  float t[NS];
  if constexpr (NS == 2) {
    const float2 v = *reinterpret_cast<const float2*>(src);
    t[0] = v.x; t[1] = v.y;
  } else {
#pragma unroll
    for (int j = 0; j < NS; j += 4) {
      const float4 v = *reinterpret_cast<const float4*>(src + j);
      t[j] = v.x; t[j + 1] = v.y; t[j + 2] = v.z; t[j + 3] = v.w;
    }
  }
#pragma unroll
  for (int j = 0; j < NS; ++j) p[j] = (double)(rev ? t[NS - 1 - j] : t[j]);
}

// x(s) <- (x(s) + x(s-1)) * p(s) in the lane's (possibly reversed) state order; `sum` keeps the pre-emission value.
// alpha: x = alpha (NoBlankCTC.py:73-85).  beta half: x(s) = beta_t(s) p_t(s), sum = beta_t(s).
// `carry` enters position 0 of the half (the virtual start state: NoBlankCTC.py:92-93 and the t>0 guard at :75).
template <int NS>
__device__ __forceinline__ void chain_step(double (&x)[NS], double (&sum)[NS], const double (&p)[NS], int hl, double& carry) {
  double up = __shfl_up_sync(0xffffffffu, x[NS - 1], 1, 16);
  if (hl == 0) up = carry;
  carry = 0.0;
#pragma unroll
  for (int j = NS - 1; j >= 1; --j) sum[j] = x[j] + x[j - 1];
  sum[0] = x[0] + up;
#pragma unroll
  for (int j = 0; j < NS; ++j) x[j] = sum[j] * p[j];
}

template <int NS>
__device__ __forceinline__ void chain_warp(const Problem& P, const FusedCfg& cfg, const Smem& S, int lane, int64_t b,
                                           int Tb, int Lb, float wgt) {
  constexpr int Lpad = 16 * NS;
  const int hl = lane & 15;
  const bool isb = lane >= 16;  // beta half (phase 2 only)
  const int NT = (Tb + kTT - 1) / kTT;
  const int NBUF = cfg.NBUFP, NBUFG = cfg.NBUFG;
  double* ck = (cfg.ckpt_global ? cfg.ws_ckpt + ((size_t)b * cfg.NTmax) * Lpad : S.ckpt) + hl;
  int* cke = cfg.ckpt_global ? cfg.ws_cke + (size_t)b * cfg.NTmax : S.cke;
  double x[NS], sum[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) x[j] = 0.0;
  int Ea = 0;
  double carry = (lane == 0) ? 1.0 : 0.0;
  PROF_DECL
  // ------------------------------------------------------------------ phase 1: alpha (lanes 16-31 carry zeros)
  for (int k = 0; k < NT; ++k) {
    const int buf = k % NBUF;
    if (k > 0) {
      Ea += rescale_half<NS>(x);
      if (!isb) {
#pragma unroll
        for (int j = 0; j < NS; ++j) ck[(k * NS + j) * 16] = x[j];
        if (lane == 0) cke[k] = Ea;
      }
    }
    PROF_SCOPE(0, mbar_wait(&S.pfull[buf], (k / NBUF) & 1))
    const float* pt = S.ptile + buf * (kTT * Lpad);
    const int nv = min(kTT, Tb - k * kTT);
    PROF_SCOPE(1,
#pragma unroll 2
    for (int i = 0; i < nv; ++i) {
      double p[NS];
      load_p<NS>(pt + i * Lpad, hl, false, p);
      chain_step<NS>(x, sum, p, hl, carry);
    })
    __syncwarp();
    if (lane == 0) count_publish(&S.cnt[0], k + 1);
  }
  // ------------------------------------------------------------------ read-out (NoBlankCTC.py:58-68,:139)
  const int sl = Lb - 1;
  double mine = 0.0;
#pragma unroll
  for (int j = 0; j < NS; ++j)
    if (j == sl % NS) mine = x[j];
  const double zhat = __shfl_sync(0xffffffffu, mine, sl / NS);
  const int Ez = __shfl_sync(0xffffffffu, Ea, 0);
  if (lane == 0) P.loss[b] = (zhat > 0.0) ? (float)(-(log(zhat) + (double)Ez * 0.6931471805599453)) : INFINITY;
  if (P.grad == nullptr) {
    PROF_DUMP(0)
    return;
  }
  const double zinv = (zhat > 0.0) ? (double)wgt / zhat : 0.0;  // sequence weight folded into gamma
  // ------------------------------------------------------------------ phase 2: beta (lanes 16-31) + alpha replay (0-15)
  // beta half: position q of the half holds state Lpad-1-q; x = u_t(s) = beta_t(s) p_t(s).  Virtual start
  // u_{T_b}(L_b) = 1 gives beta_{T_b-1}(L_b-1) = 1 without a branch (state L_b has p = 0 and alpha = 0).
  int Eb = 0;
  if (isb) {
#pragma unroll
    for (int j = 0; j < NS; ++j) x[j] = (Lpad - 1 - (hl * NS + j) == Lb) ? 1.0 : 0.0;
    carry = (hl == 0 && Lb == Lpad) ? 1.0 : 0.0;
  }
  double* at = S.abtile;               // [kTT][Lpad] alpha_t(s) (pre-scaled)
  double* bt = S.abtile + kTT * Lpad;  // [kTT][Lpad] beta_t(s)
  for (int j2 = 0; j2 < NT; ++j2) {
    const int k = NT - 1 - j2;
    const int n = NT + j2;
    const int buf = n % NBUF;
    const int gbuf = j2 % NBUFG;
    // gamma = alpha * beta * w / Z; the power-of-two part is split over both factors (range safety)
    const int EaK = (k == 0) ? 0 : cke[k];
    const int Eb_all = __shfl_sync(0xffffffffu, Eb, 16);
    const int d = EaK + Eb_all - Ez;
    const double s1 = pow2i(d / 2);
    const double s2 = -(pow2i(d - d / 2) * zinv);  // negative: the row warps ADD gamma' = -w*gamma to the softmax row
    if (!isb) {
      if (k == 0) {
#pragma unroll
        for (int j = 0; j < NS; ++j) x[j] = 0.0;
        carry = (lane == 0) ? s1 : 0.0;
      } else {
#pragma unroll
        for (int j = 0; j < NS; ++j) x[j] = ck[(k * NS + j) * 16] * s1;  // exact: alpha replay runs pre-scaled
      }
    }
    PROF_SCOPE(2, mbar_wait(&S.pfull[buf], (n / NBUF) & 1))
    const float* pt = S.ptile + buf * (kTT * Lpad);
    const int nv = min(kTT, Tb - k * kTT);
    // position -> state index of this lane's slots
    const int s0 = isb ? (Lpad - 1 - hl * NS) : hl * NS;
    const int sdir = isb ? -1 : 1;
    double* dst = (isb ? bt : at) + s0;
    PROF_SCOPE(3,
#pragma unroll 2
    for (int jj = 0; jj < nv; ++jj) {
      const int i = isb ? (nv - 1 - jj) : jj;  // alpha walks up the tile, beta walks down
      double p[NS];
      load_p<NS>(pt + i * Lpad, hl, isb, p);
      chain_step<NS>(x, sum, p, hl, carry);
#pragma unroll
      for (int j = 0; j < NS; ++j) dst[i * Lpad + sdir * j] = isb ? sum[j] : x[j];
    })
    if (j2 >= NBUFG) PROF_SCOPE(4, mbar_wait(&S.gempty[gbuf], ((j2 / NBUFG) - 1) & 1))
    __syncwarp();
    float* gt = S.gtile + gbuf * (kTT * Lpad);
    PROF_SCOPE(5, for (int idx = lane; idx < nv * Lpad; idx += 32) gt[idx] = (float)(at[idx] * (bt[idx] * s2));)
    {
      // every lane takes part in the half-wide shuffles; only the beta half keeps the result
      const int e = rescale_half<NS>(x);
      if (isb) Eb += e;
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&S.gfull[gbuf]);
      count_publish(&S.cnt[0], n + 1);
    }
  }
  PROF_DUMP(0)
}

// ============================================================================ TMA producer warp
__device__ __forceinline__ void producer_warp(const Problem& P, const FusedCfg& cfg, const Smem& S, int lane, int64_t b,
                                              int Tb) {
  if (lane != 0) return;
  const int NT = (Tb + kTT - 1) / kTT;
  const int total = (P.grad != nullptr) ? 2 * NT : NT;
  const uint32_t C4 = (uint32_t)P.C * 4u;
  const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
  const uint64_t stride = (uint64_t)P.B * C4;  // bytes between rows t and t+1 of one sequence
  const uint64_t seq0 = reinterpret_cast<uint64_t>(P.logits) + (uint64_t)b * C4;
  const uint64_t end = reinterpret_cast<uint64_t>(cfg.logits_end);
  const bool guard_last = (end & 15) != 0 && b == P.B - 1 && Tb == P.T;  // only the tensor's last row can over-read
  const uint32_t ring0 = smem_u32(S.ring);
  const uint32_t slot_bytes = (uint32_t)kTT * cfg.RS;
  int slot = 0, use = 0;
  PROF_DECL
  for (int n = 0; n < total; ++n) {
    const int k = n < NT ? n : 2 * NT - 1 - n;
    if (use >= 1) PROF_SCOPE(n < NT ? 0 : 1, mbar_wait(&S.sempty[slot], (use - 1) & 1))
    const int nv = min(kTT, Tb - k * kTT);
    const uint32_t bar = smem_u32(&S.sfull[slot]);
    const uint64_t pol = n < NT ? pol_keep : pol_stream;
    uint32_t dst = ring0 + slot * slot_bytes;
    uint64_t a = seq0 + (uint64_t)k * kTT * stride;
    uint32_t bytes = 0;
    for (int i = 0; i < nv; ++i, a += stride, dst += cfg.RS) {
      const uint32_t nb = (((uint32_t)a & 15u) + C4 + 15u) & ~15u;
      if (guard_last && k * kTT + i == P.T - 1) {
        // the very last row of an unaligned tensor would be over-read by < 16 bytes: copy it by hand
        const float* src = reinterpret_cast<const float*>(a);
        float* d = reinterpret_cast<float*>(S.ring + (size_t)slot * slot_bytes + (size_t)i * cfg.RS) + ((a >> 2) & 3);
        for (uint32_t c = 0; c < (uint32_t)P.C; ++c) d[c] = __ldg(src + c);
      } else {
        bulk_g2s(dst, a & ~uint64_t(15), nb, bar, pol);
        bytes += nb;
      }
    }
    // the phase cannot complete before this arrival, so expecting the bytes after issuing the copies is safe
    mbar_arrive_expect_tx(&S.sfull[slot], bytes);
    count_publish(&S.cnt[1], n + 1);
    if (++slot == cfg.NSLOT) { slot = 0; ++use; }
  }
  PROF_DUMP(5)
}

// ============================================================================ row warps
// Geometry of one (t,b) row seen as 16-byte chunks: the row starts `off4` floats into chunk 0 and ends `rem`
// floats into chunk nch-1 (rows are only 4-byte aligned when C % 4 != 0).  `srow` is the row's copy in the
// shared-memory ring (same 16-byte phase as in global memory).
struct RowGeom {
  float4* srow;
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
  static constexpr int Lpad = 16 * NS;
  static constexpr int NSL = Lpad / LPR;  // states per lane in the emission gather / gamma scatter
  static constexpr bool kLabRegs = NSL <= 8;
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
  int labr[kLabRegs ? NSL : 1];  // this lane's labels (-1 = no state)

  __device__ __forceinline__ Rows(const Problem& P_, const FusedCfg& cfg_, const Smem& S_, int lane_, int wrow_,
                                  int64_t b_, int Tb_, int Lb_, float wgt_)
      : P(P_), cfg(cfg_), S(S_), lane(lane_), li(lane_ & (LPR - 1)), gi(lane_ / LPR), wrow(wrow_), b(b_), Tb(Tb_),
        Lb(Lb_), C((int)P_.C), lse_arr(cfg_.lse_global ? cfg_.ws_lse + (size_t)b_ * P_.T : S_.lse),
        pol_stream(policy_evict_first()), wgt(wgt_), base_addr(reinterpret_cast<uintptr_t>(P_.logits)) {
    if constexpr (kLabRegs) {
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const int st = li + j * LPR;
        labr[j] = st < Lb_ ? S_.lab[st] : -1;
      }
    }
  }
  __device__ __forceinline__ int label(int j) const {
    if constexpr (kLabRegs) return labr[j];
    const int st = li + j * LPR;
    return st < Lb ? S.lab[st] : -1;
  }

  // row t of this sequence; `slot_rows` = start of the ring slot holding the tile, i = row inside the tile
  __device__ __forceinline__ RowGeom geom(int t, unsigned char* slot_rows, int i) const {
    RowGeom g;
    g.goff = ((int64_t)t * P.B + b) * C;
    g.off4 = (int)(((base_addr >> 2) + (uintptr_t)g.goff) & 3);
    g.nch = (g.off4 + C + 3) >> 2;
    g.rem = g.off4 + C - 4 * (g.nch - 1);
    g.srow = reinterpret_cast<float4*>(slot_rows + (size_t)i * cfg.RS);
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
      float xv[NSL];
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const int l = label(j);
        xv[j] = xr[l >= 0 ? l : 0];
      }
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const float pv = label(j) >= 0 ? fmaxf(ex2f(fmaf(xv[j], kLog2e, -lb2)), kPMin) : 0.f;
        ptile_buf[i * Lpad + li + j * LPR] = pv;
      }
    }
  }

  // ---------------------------------------------------------------- phase 1: one tile
  __device__ __forceinline__ void forward_tile(int k, int nv, unsigned char* slot_rows, float* ptile_buf) const {
#pragma unroll 1
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
  __device__ __forceinline__ void emit_tile(int k, int nv, unsigned char* slot_rows, float* ptile_buf) const {
#pragma unroll 1
    for (int pr = 0; pr < NP; ++pr) {
      const int i = pr * R + gi;
      const bool act = i < nv;
      const RowGeom g = geom(k * kTT + i, slot_rows, i);
      emit_row(g, act, i, act ? lse_arr[k * kTT + i] : 0.f, ptile_buf);
    }
  }

  // ---------------------------------------------------------------- gradient row pieces
  // chunks of one row segment -> global; only chunk 0 and chunk nch-1 can be partial
  template <bool kSingle>
  __device__ __forceinline__ void store_seg(float* grow, const RowGeom& g, int seg, const float4 (&v)[CPL]) const {
    float4* dst = reinterpret_cast<float4*>(grow - g.off4) + seg * SEG + li;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int q = seg * SEG + li + c * LPR;
      if (q < g.nch) {
        const bool may_partial = !kSingle || c == 0 || c + 2 >= CPL;
        const bool head = (q == 0) && g.off4 != 0;
        const bool tail = (q == g.nch - 1) && g.rem != 4;
        if (!may_partial || (!head && !tail)) {
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
  // in place in the row's shared-memory copy: x -> w * softmax(x)
  __device__ __forceinline__ void softmax_seg(const RowGeom& g, int seg, float lb2) const {
    float4* src = g.srow + seg * SEG + li;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int q = seg * SEG + li + c * LPR;
      if (q < g.nch) {
        float4 x = src[c * LPR];
        x.x = wgt * ex2f(fmaf(x.x, kLog2e, -lb2));
        x.y = wgt * ex2f(fmaf(x.y, kLog2e, -lb2));
        x.z = wgt * ex2f(fmaf(x.z, kLog2e, -lb2));
        x.w = wgt * ex2f(fmaf(x.w, kLog2e, -lb2));
        src[c * LPR] = x;
      }
    }
  }
  template <bool kSingle>
  __device__ __forceinline__ void copy_out_seg(float* grow, const RowGeom& g, int seg) const {
    const float4* src = g.srow + seg * SEG + li;
    float4 v[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int q = seg * SEG + li + c * LPR;
      if (q < g.nch) v[c] = src[c * LPR];
    }
    store_seg<kSingle>(grow, g, seg, v);
  }

  // ---------------------------------------------------------------- phase 2 stage B: one tile
  // w*softmax is formed in place in the row's shared-memory copy, -w*gamma is scattered onto it with shared-memory
  // atomics (repeated labels accumulate), then the row streams out with 128-bit stores.
  __device__ __forceinline__ void backward_tile(int k, int nv, unsigned char* slot_rows, const float* gtile_buf) const {
#pragma unroll 1
    for (int pr = 0; pr < NP; ++pr) {
      const int i = pr * R + gi;
      const bool act = i < nv;
      const RowGeom g = geom(k * kTT + i, slot_rows, i);
      if (act) {
        const float lb2 = lse_arr[k * kTT + i] * kLog2e;
        for (int seg = 0; seg < cfg.NSEG; ++seg) softmax_seg(g, seg, lb2);
      }
      __syncwarp();
      if (act) {
        float* xr = reinterpret_cast<float*>(g.srow) + g.off4;
#pragma unroll
        for (int j = 0; j < NSL; ++j) {
          const int l = label(j);
          if (l >= 0) atomicAdd(&xr[l], gtile_buf[i * Lpad + li + j * LPR]);
        }
      }
      __syncwarp();
      if (act) {
        float* grow = P.grad + g.goff;
        if (cfg.NSEG == 1) {
          copy_out_seg<true>(grow, g, 0);
        } else {
          for (int seg = 0; seg < cfg.NSEG; ++seg) copy_out_seg<false>(grow, g, seg);
        }
      }
    }
  }

  // rows t in [t_begin, T) get an all-zero gradient (grads beyond input_length are exactly 0)
  __device__ __forceinline__ void zero_rows(int t_begin) const {
    float4 z[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) z[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t t = (int64_t)t_begin + wrow * R + gi; t < P.T; t += kNW * R) {
      const RowGeom g = geom((int)t, nullptr, 0);
      for (int seg = 0; seg < cfg.NSEG; ++seg) store_seg<false>(P.grad + g.goff, g, seg, z);
    }
  }

  __device__ __forceinline__ void run() const {
    const int NT = (Tb + kTT - 1) / kTT;
    const int NBUF = cfg.NBUFP, NBUFG = cfg.NBUFG, NSLOT = cfg.NSLOT;
    const size_t slot_bytes = (size_t)kTT * cfg.RS;
    PROF_DECL
    // ---- phase 1: tiles k = wrow, wrow+NW, ...
    for (int k = wrow; k < NT; k += kNW) {
      const int buf = k % NBUF;
      const int slot = k % NSLOT;
      PROF_SCOPE(0,
      if (k >= NBUF) count_wait(&S.cnt[0], k - NBUF + 1);  // chain is done with the p-tile buffer's previous use
      count_wait(&S.cnt[1], k + 1);                         // the copy has been issued => the parity wait is sound
      mbar_wait(&S.sfull[slot], (k / NSLOT) & 1);)
      PROF_SCOPE(1, forward_tile(k, min(kTT, Tb - k * kTT), S.ring + slot * slot_bytes, S.ptile + buf * (kTT * Lpad));)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&S.pfull[buf]);
        mbar_arrive(&S.sempty[slot]);
      }
    }
    if (P.grad == nullptr) {
      PROF_DUMP(1 + wrow)
      return;
    }
    if (Tb < P.T) zero_rows(Tb);
    // ---- phase 2: the same warp owns the same tiles (it wrote their lse values), walked downwards:
    // A(k) = emissions -> chain, B(k) = gradient rows once the chain has produced gamma(k).
    int k = NT - 1 - ((NT - 1 - wrow) % kNW + kNW) % kNW;  // largest k <= NT-1 with k % NW == wrow
    for (; k >= 0; k -= kNW) {
      const int j2 = NT - 1 - k;
      const int n = NT + j2;
      const int buf = n % NBUF, gbuf = j2 % NBUFG, slot = n % NSLOT;
      const int nv = min(kTT, Tb - k * kTT);
      unsigned char* rows = S.ring + slot * slot_bytes;
      PROF_SCOPE(2,
      if (n >= NBUF) count_wait(&S.cnt[0], n - NBUF + 1);
      count_wait(&S.cnt[1], n + 1);
      mbar_wait(&S.sfull[slot], (n / NSLOT) & 1);)
      PROF_SCOPE(3, emit_tile(k, nv, rows, S.ptile + buf * (kTT * Lpad));)
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.pfull[buf]);
      PROF_SCOPE(4, mbar_wait(&S.gfull[gbuf], (j2 / NBUFG) & 1))
      PROF_SCOPE(5, backward_tile(k, nv, rows, S.gtile + gbuf * (kTT * Lpad));)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&S.gempty[gbuf]);
        mbar_arrive(&S.sempty[slot]);
      }
    }
    PROF_DUMP(1 + wrow)
  }
};

// ============================================================================ kernel
template <int NS, int LPR, int CPL>
__global__ void __launch_bounds__(kThreads, (NS <= 4 ? 4 : 2)) nbctc_fused_kernel(const Problem P, const FusedCfg cfg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem S;
  S.pfull = reinterpret_cast<uint64_t*>(smem_raw + cfg.o_bar);
  S.gfull = S.pfull + kMaxBuf;
  S.gempty = S.gfull + kMaxBuf;
  S.sfull = S.gempty + kMaxBuf;
  S.sempty = S.sfull + kMaxSlot;
  S.cnt = reinterpret_cast<volatile int*>(S.sempty + kMaxSlot);
  S.lab = reinterpret_cast<int*>(smem_raw + cfg.o_lab);
  S.lse = reinterpret_cast<float*>(smem_raw + cfg.o_lse);
  S.ckpt = reinterpret_cast<double*>(smem_raw + cfg.o_ckpt);
  S.cke = reinterpret_cast<int*>(smem_raw + cfg.o_cke);
  S.ptile = reinterpret_cast<float*>(smem_raw + cfg.o_ptile);
  S.gtile = reinterpret_cast<float*>(smem_raw + cfg.o_gtile);
  S.abtile = reinterpret_cast<double*>(smem_raw + cfg.o_abtile);
  S.ring = smem_raw + cfg.o_ring;

  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t Tb64 = P.in_len[b], Lb64 = P.tgt_len[b];
  bool ok = seq_feasible(Tb64, Lb64, P.T, P.Lmax);
  const int Tb = (int)Tb64, Lb = (int)Lb64;
  int bad = 0;
  if (ok) {
    for (int s = tid; s < 16 * NS; s += blockDim.x) {
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
    mbar_init(&S.gfull[tid], 1);
    mbar_init(&S.gempty[tid], 1);
  }
  if (tid < kMaxSlot) {
    mbar_init(&S.sfull[tid], 1);
    mbar_init(&S.sempty[tid], 1);
  }
  if (tid < 2) S.cnt[tid] = 0;
  static_assert(kNW == 4, "the gamma-tile ring depth must equal the number of row warps (same waiter per buffer)");
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
