// Block-streaming no-blank CTC forward+backward kernel for sm_100a (overview in nbctc_stream.cu).
//
// One CTA owns a GROUP of GB = 32/LPR batch-adjacent sequences and walks their lattice in tiles of TT time steps.
// The rows (t, b0..b0+GB) of one time step are contiguous in the (T,B,C) tensor ("slab"), so a tile is TT bulk
// copies of GB*C*4 bytes.  The CTA runs a lock-step software pipeline over "items" (item i < NT: phase-1 tile i,
// walking up; item i >= NT: phase-2 tile 2NT-1-i, walking down), one __syncthreads() per iteration, no polling:
//
//   iteration `it`:   producer warp     lane i moves time step i of a tile: TMA bulk store of item it-2's gradient
//                                       rows; TMA bulk loads of the items ahead (as far as ring slots are free)
//                                       with an L2 evict_last policy in phase 1 (phase 2 re-reads hit L2)
//                     row warps         warp w = time step w of the tile, lane group = sequence.  Item it+1: row
//                                       log-partition + emission gather (phase 1, NoBlankCTC.py:136,:96-102) or
//                                       emission gather only (phase 2); item it-1: w*softmax over the whole slab
//                                       in place in the ring slot, then scatter of -w*gamma
//                     chain warps       item it: one warp per sequence, 16 lanes x NS states in float64, linear
//                                       domain with exact power-of-two rescaling per tile (NoBlankCTC.py:71-87).
//                                       Phase 1: alpha + one checkpoint per tile.  Phase 2: lanes 0-15 replay
//                                       alpha inside the tile from the checkpoint while lanes 16-31 run beta in
//                                       the same instructions (beta is kept in reversed state order).
//
// Template parameters: NS (chain states per lane, Lmax <= 16*NS), LPR (lanes per row; GB = 32/LPR sequences per
// CTA), CPL (16-byte chunks per lane and row segment).
#pragma once

#include <cuda_runtime.h>

#include "common.cuh"

namespace nbctc {

constexpr int kMaxGB = 8;   // sequences per group upper bound (LPR = 4)
constexpr int kNSlot = 6;   // ring depth (tiles)

struct StreamCfg {
  int NS, Lpad, TT;
  int LPR, CPL, NSEG;
  int GB;     // sequences per group (CTA) = 32 / LPR
  int RSg;    // bytes per time step in a ring slot: round16(GB*C*4) + 32
  int NTmax;  // ceil(T / TT)
  int Tpad;   // lse row stride (floats)
  int ckpt_global, lse_global;
  int ctas_per_sm;
  uint32_t o_bar, o_info, o_lab, o_lse, o_ckpt, o_cke, o_ptile, o_ab, o_s2, o_tab, o_ring, smem_bytes;
  double* ws_ckpt;  // [B][NTmax][Lpad]
  int* ws_cke;      // [B][NTmax]
  float* ws_lse;    // [B][T]
  long long* prof;  // role profiler output (NBCTC_PROF builds), else null: [3][8] buckets, then [128][32][2] trace
};

template <int NS, int LPR>
struct Geo {
  static constexpr int TT = NS >= 16 ? 4 : 8;  // time steps per tile
  static constexpr int Lpad = 16 * NS;
  static constexpr int GB = 32 / LPR;          // sequences per CTA
  static constexpr int NRW = TT;               // row warps: one per time step of a tile
  static constexpr int NSL = Lpad / LPR;       // states per lane in the emission gather / gamma scatter
  static constexpr int PS = Lpad + 8;          // p-tile row stride (floats)
  static constexpr int AS = Lpad + 8;          // alpha/beta tile row stride (doubles)
  static constexpr int PSEQ = TT * PS + 8;     // p-tile floats per sequence (+8: lane groups hit distinct banks)
  static constexpr int ABSEQ = 2 * TT * AS + 8;  // alpha+beta tile doubles per sequence
  static constexpr int NTHREADS = 32 * (GB + NRW + 1);
};

int launch_stream_ns2(const Problem& p, const StreamCfg& cfg, cudaStream_t stream);
int launch_stream_ns4(const Problem& p, const StreamCfg& cfg, cudaStream_t stream);
int launch_stream_ns8(const Problem& p, const StreamCfg& cfg, cudaStream_t stream);
int launch_stream_ns16(const Problem& p, const StreamCfg& cfg, cudaStream_t stream);

#ifdef __CUDACC__
namespace stream {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kPMin = 7.52316385e-37f;  // 2^-120: emission floor (8 steps stay inside the f64 range)
constexpr float kNegInf = -INFINITY;
// label slots in shared memory / registers: class index in the low 22 bits, above it the state's rank among the
// earlier states with the same class (the gamma scatter runs one conflict-free round per rank); -1 = no state
constexpr int kLabBits = 22;
constexpr int kLabMask = (1 << kLabBits) - 1;

// ---------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
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
// try_wait suspends the warp in hardware until the phase completes or a time limit passes (no hot spin).  A
// protocol bug would hang the GPU: after ~2 s the kernel traps instead, which a correct run never reaches.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t bar = smem_u32(b);
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
// TMA bulk copies (1-D, 16-byte aligned, size a multiple of 16)
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, uint64_t src_gmem, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          dst_smem),
      "l"(src_gmem), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, uint64_t src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(uint64_t dst_gmem, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
__device__ __forceinline__ double pow2i(int e) {  // exact 2^e, e clamped to the normal range
  e = max(-1022, min(1023, e));
  return __hiloint2double((1023 + e) << 20, 0);
}

// Optional role profiler (compile with -DNBCTC_PROF): per-warp cycle buckets, compiled out of the product build.
#ifdef NBCTC_PROF
#define PROF_DECL long long prof_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long prof_t0_ = clock64();
#define PROF_SCOPE(i, ...) { const long long t_ = clock64(); __VA_ARGS__; prof_[i] += clock64() - t_; }
#define PROF_DUMP(role_)                                                                                     \
  if ((threadIdx.x & 31) == 0 && cfg.prof != nullptr) {                                                      \
    prof_[7] = clock64() - prof_t0_;                                                                         \
    for (int i_ = 0; i_ < 8; ++i_) atomicAdd((unsigned long long*)&cfg.prof[(role_) * 8 + i_], (unsigned long long)prof_[i_]); \
  }
#else
#define PROF_DECL
#define PROF_SCOPE(i, ...) { __VA_ARGS__; }
#define PROF_DUMP(role_)
#endif

struct Smem {
  uint64_t* sfull;  // [kNSlot] ring-slot "rows have landed" barriers
  int* info;        // [0..GB) T_b, [kMaxGB..) L_b, [2*kMaxGB..) bad-label flags, [3*kMaxGB..) largest duplicate rank
  int* lab;         // [GB][Lpad]
  float* lse;       // [GB][Tpad]
  double* ckpt;     // [GB][NTmax][Lpad]
  int* cke;         // [GB][NTmax]
  float* ptile;     // [2][GB][PSEQ]
  double* ab;       // [2][GB][ABSEQ]
  double* s2;       // [2][GB]
  float2* tab;      // [NRW][kMaxGB + 1] per row warp: (lse*log2e, w) of the slab's rows
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

// emissions of the lane's NS states for one row of a p-tile, in the lane's own state order: the beta half
// (rev) holds state Lpad-1-q at position q
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

// Chain-warp state lives in plain registers of the kernel body (passed by reference to force-inlined functions).
struct ChainScal {
  double carry, zinv;
  int Ea, Eb, Ez;
};

// ---- phase 1, tile k: alpha over the tile's steps (lanes 16-31 carry zeros)
template <int NS, int TT, int PS>
__device__ __forceinline__ void chain_phase1(double (&x)[NS], ChainScal& c, int lane, int Tb, double* ck, int* cke, int k,
                                             const float* __restrict__ pt) {
  const int hl = lane & 15;
  double sum[NS];
  if (k > 0) {
    c.Ea += rescale_half<NS>(x);
    if (lane < 16) {
#pragma unroll
      for (int j = 0; j < NS; ++j) ck[(k * NS + j) * 16] = x[j];
      if (lane == 0) cke[k] = c.Ea;
    }
  }
  const int nv = min(TT, Tb - k * TT);
  if (NS <= 4 && nv == TT) {
    // whole tile of emissions in registers ahead of the dependent loop
    double pr[TT][NS];
#pragma unroll
    for (int i = 0; i < TT; ++i) load_p<NS>(pt + i * PS, hl, false, pr[i]);
#pragma unroll
    for (int i = 0; i < TT; ++i) chain_step<NS>(x, sum, pr[i], hl, c.carry);
  } else {
#pragma unroll 2
    for (int i = 0; i < nv; ++i) {
      double pf[NS];
      load_p<NS>(pt + i * PS, hl, false, pf);
      chain_step<NS>(x, sum, pf, hl, c.carry);
    }
  }
}

// ---- read-out after the sequence's last phase-1 tile (NoBlankCTC.py:58-68,:139) + beta start state
template <int NS>
__device__ __forceinline__ void chain_readout(double (&x)[NS], ChainScal& c, int lane, int Lb, float* loss_out, float wgt) {
  constexpr int Lpad = 16 * NS;
  const int hl = lane & 15;
  const int sl = Lb - 1;
  // x[sl % NS] without a dynamic index (which would push the state array into local memory)
  const int rj = sl % NS;
  double mine = x[0];
#pragma unroll
  for (int j = 1; j < NS; ++j) mine = (rj >= j) ? x[j] : mine;
  const double zhat = __shfl_sync(0xffffffffu, mine, sl / NS);
  c.Ez = __shfl_sync(0xffffffffu, c.Ea, 0);
  if (lane == 0) *loss_out = (zhat > 0.0) ? (float)(-(log(zhat) + (double)c.Ez * 0.6931471805599453)) : INFINITY;
  c.zinv = (zhat > 0.0) ? (double)wgt / zhat : 0.0;  // sequence weight folded into gamma
  // beta half: position q of the half holds state Lpad-1-q; x = u_t(s) = beta_t(s) p_t(s).  Virtual start
  // u_{T_b}(L_b) = 1 gives beta_{T_b-1}(L_b-1) = 1 without a branch (state L_b has p = 0 and alpha = 0).
  c.Eb = 0;
  if (lane >= 16) {
#pragma unroll
    for (int j = 0; j < NS; ++j) x[j] = (Lpad - 1 - (hl * NS + j) == Lb) ? 1.0 : 0.0;
    c.carry = (hl == 0 && Lb == Lpad) ? 1.0 : 0.0;
  }
}

// ---- phase 2, tile k: beta (lanes 16-31) + alpha replay (lanes 0-15); alpha_t(s), beta_t(s) -> ab tile
template <int NS, int TT, int PS, int AS>
__device__ __forceinline__ void chain_phase2(double (&x)[NS], ChainScal& c, int lane, int Tb, const double* ck, const int* cke,
                                             int k, const float* __restrict__ pt, double* __restrict__ abt, double* s2_out) {
  constexpr int Lpad = 16 * NS;
  const int hl = lane & 15;
  const bool isb = lane >= 16;
  double sum[NS];
  // gamma = alpha * beta * w / Z; the power-of-two part is split over both factors (range safety)
  const int EaK = (k == 0) ? 0 : cke[k];
  const int Eb_all = __shfl_sync(0xffffffffu, c.Eb, 16);
  const int d = EaK + Eb_all - c.Ez;
  const double s1 = pow2i(d / 2);
  const double s2 = -(pow2i(d - d / 2) * c.zinv);  // negative: the row warps ADD gamma' = -w*gamma to the softmax row
  if (lane == 0) *s2_out = s2;
  if (!isb) {
    if (k == 0) {
#pragma unroll
      for (int j = 0; j < NS; ++j) x[j] = 0.0;
      c.carry = (lane == 0) ? s1 : 0.0;
    } else {
#pragma unroll
      for (int j = 0; j < NS; ++j) x[j] = ck[(k * NS + j) * 16] * s1;  // exact: alpha replay runs pre-scaled
    }
  }
  const int nv = min(TT, Tb - k * TT);
  // position -> state index of this lane's slots
  const int s0 = isb ? (Lpad - 1 - hl * NS) : hl * NS;
  const int sdir = isb ? -1 : 1;
  double* dst = abt + (isb ? TT * AS : 0) + s0;  // alpha tile, then beta tile
  if (NS <= 4 && nv == TT) {
    double pr[TT][NS];
#pragma unroll
    for (int jj = 0; jj < TT; ++jj) load_p<NS>(pt + (isb ? TT - 1 - jj : jj) * PS, hl, isb, pr[jj]);
#pragma unroll
    for (int jj = 0; jj < TT; ++jj) {
      const int i = isb ? (TT - 1 - jj) : jj;  // alpha walks up the tile, beta walks down
      chain_step<NS>(x, sum, pr[jj], hl, c.carry);
#pragma unroll
      for (int j = 0; j < NS; ++j) dst[i * AS + sdir * j] = isb ? sum[j] : x[j];
    }
  } else {
#pragma unroll 2
    for (int jj = 0; jj < nv; ++jj) {
      const int i = isb ? (nv - 1 - jj) : jj;
      double pf[NS];
      load_p<NS>(pt + i * PS, hl, isb, pf);
      chain_step<NS>(x, sum, pf, hl, c.carry);
#pragma unroll
      for (int j = 0; j < NS; ++j) dst[i * AS + sdir * j] = isb ? sum[j] : x[j];
    }
  }
  // every lane takes part in the half-wide shuffles; only the beta half keeps the result
  const int e = rescale_half<NS>(x);
  if (isb) c.Eb += e;
}

// ============================================================================ row warps
// Geometry of one (t,b) row seen as 16-byte chunks of the ring slot: the row starts `off4` floats into chunk 0
// and ends `rem` floats into chunk nch-1 (rows are only 4-byte aligned when C % 4 != 0, and neighbouring
// sequences of the group share their boundary chunks).
struct RowGeom {
  float4* srow;
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

// One row warp = one time step of every tile; lane group gi (LPR lanes) = sequence gi of the CTA's group.
template <int NS, int LPR, int CPL>
struct Rows {
  using G = Geo<NS, LPR>;
  static constexpr int TT = G::TT, Lpad = G::Lpad, PS = G::PS, AS = G::AS, GB = G::GB, NSL = G::NSL;
  static constexpr bool kLabRegs = NSL <= 8;
  static constexpr int SEG = LPR * CPL;  // chunks per row segment

  const Problem& P;
  const StreamCfg& cfg;
  const int lane, li, seq;  // seq = lane group = sequence of the group
  const int ti;             // this warp's time step inside a tile
  const int gcnt;
  const int64_t b0;
  const int Tb, Lb, C;
  const int max_rank;
  float* lse_seq;
  const int* lab_seq;
  float2* tab;
  const int* info;
  int labr[kLabRegs ? NSL : 1];  // this lane's labels (-1 = no state)

  __device__ __forceinline__ Rows(const Problem& P_, const StreamCfg& cfg_, const Smem& S_, int lane_, int ti_, int gcnt_,
                                  int64_t b0_)
      : P(P_), cfg(cfg_), lane(lane_), li(lane_ & (LPR - 1)), seq(lane_ / LPR), ti(ti_), gcnt(gcnt_), b0(b0_),
        Tb(S_.info[lane_ / LPR]), Lb(S_.info[kMaxGB + lane_ / LPR]), C((int)P_.C), max_rank(S_.info[3 * kMaxGB + lane_ / LPR]),
        lse_seq(cfg_.lse_global ? cfg_.ws_lse + (size_t)min(b0_ + lane_ / LPR, P_.B - 1) * P_.T
                                : S_.lse + (size_t)(lane_ / LPR) * cfg_.Tpad),
        lab_seq(S_.lab + (lane_ / LPR) * Lpad), tab(S_.tab + ti_ * (kMaxGB + 1)), info(S_.info) {
    if constexpr (kLabRegs) {
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const int st = li + j * LPR;
        labr[j] = st < Lb ? lab_seq[st] : -1;
      }
    }
  }
  __device__ __forceinline__ int label(int j) const {
    if constexpr (kLabRegs) return labr[j];
    const int st = li + j * LPR;
    return st < Lb ? lab_seq[st] : -1;
  }

  // float index (0..3) of the slab's first element inside its first 16-byte chunk (both tensors are 16-byte aligned)
  __device__ __forceinline__ int slab_phase(int t) const {
    return (int)(((((unsigned)t & 3u) * ((unsigned)P.B & 3u) + ((unsigned)b0 & 3u)) * ((unsigned)C & 3u)) & 3u);
  }
  // this lane group's row at time t; `tsl` = the time step's slab in the ring slot
  __device__ __forceinline__ RowGeom geom(int t, unsigned char* tsl) const {
    const int fidx = slab_phase(t) + seq * C;  // float index of the row inside the slab's chunks
    RowGeom g;
    g.off4 = fidx & 3;
    g.nch = (g.off4 + C + 3) >> 2;
    g.rem = g.off4 + C - 4 * (g.nch - 1);
    g.srow = reinterpret_cast<float4*>(tsl) + (fidx >> 2);
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
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        s0 += ex2f(fmaf(v[c].x, kLog2e, -mb));
        s1 += ex2f(fmaf(v[c].y, kLog2e, -mb));
        s2 += ex2f(fmaf(v[c].z, kLog2e, -mb));
        s3 += ex2f(fmaf(v[c].w, kLog2e, -mb));
      }
      s_run = (m_run > kNegInf ? s_run * ex2f((m_run - m) * kLog2e) : 0.f) + ((s0 + s1) + (s2 + s3));
      m_run = m;
    }
  }

  // emissions p_t(s) = softmax(x_t)[label_s] (gathered from the row's shared-memory copy) -> p-tile row
  __device__ __forceinline__ void emit_row(const RowGeom& g, bool act, float lse, float* prow) const {
    if (act) {
      const float* xr = reinterpret_cast<const float*>(g.srow) + g.off4;
      const float lb2 = lse * kLog2e;
      float xv[NSL];
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const int l = label(j);
        xv[j] = xr[l >= 0 ? (l & kLabMask) : 0];
      }
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const float pv = label(j) >= 0 ? fmaxf(ex2f(fmaf(xv[j], kLog2e, -lb2)), kPMin) : 0.f;
        prow[li + j * LPR] = pv;
      }
    }
  }

  // ---------------------------------------------------------------- phase 1: row log-partition + emissions
  // tsl: this warp's time step slab of the tile; pt: p-tile of the item ([GB][PSEQ])
  __device__ __forceinline__ void forward_step(int t, unsigned char* tsl, float* pt) const {
    const bool act = t < Tb;  // Tb = 0 for sequences outside the group / the parity domain
    const RowGeom g = geom(t, tsl);
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
    if (act && li == 0) lse_seq[t] = lse;
    emit_row(g, act, lse, pt + seq * G::PSEQ + ti * PS);
  }

  // ---------------------------------------------------------------- phase 2 ahead stage: emissions again
  __device__ __forceinline__ void emit_step(int t, unsigned char* tsl, float* pt) const {
    const bool act = t < Tb;
    const RowGeom g = geom(t, tsl);
    emit_row(g, act, act ? lse_seq[t] : 0.f, pt + seq * G::PSEQ + ti * PS);
  }

  // w * softmax element; rc = (lse*log2e, w) of the element's row.  w = 0 (row beyond input_length, or stale
  // ring contents that were never loaded) gives an exact 0 whatever x holds
  static __device__ __forceinline__ float soft1(float x, float2 rc) {
    return rc.y != 0.f ? rc.y * ex2f(fmaf(x, kLog2e, -rc.x)) : 0.f;
  }

  // ---------------------------------------------------------------- phase 2 behind stage: gradient rows
  // The whole slab (GB rows, contiguous) becomes w*softmax(x) in place in the ring slot, chunk by chunk with
  // the row constants of the (at most two) rows a chunk touches: zeros beyond input_length.
  // Then -w*gamma is scattered onto each lane group's row; the producer warp streams the slab out with one
  // TMA bulk store.  abt: alpha/beta tiles of the item ([GB][ABSEQ]); s2v: per-sequence gamma scale.
  __device__ __forceinline__ void grad_step(int t, unsigned char* tsl, const double* abt, const double* s2v) const {
    // row constants of the slab -> per-warp table (entry GB = "no row": exp2(-inf) = 0)
    if (lane <= GB) {
      const bool live_r = lane < GB && t < info[lane];
      const float w_r = live_r ? P.w_scalar * (P.seq_w ? P.seq_w[b0 + lane] : 1.f) : 0.f;
      const float* lse_r = cfg.lse_global ? cfg.ws_lse + (size_t)min(b0 + lane, P.B - 1) * P.T : lse_seq + (lane - seq) * cfg.Tpad;
      tab[lane] = make_float2(live_r ? lse_r[t] * kLog2e : INFINITY, w_r);
    }
    __syncwarp();
    const int ph = slab_phase(t);
    const int nchs = (ph + gcnt * C + 3) >> 2;  // chunks of the slab
    float4* slab4 = reinterpret_cast<float4*>(tsl);
    for (int seg = 0; seg < cfg.NSEG; ++seg) {
      float4 v[CPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int q = (seg * CPL + c) * 32 + lane;
        if (q < nchs) v[c] = slab4[q];
      }
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int q = (seg * CPL + c) * 32 + lane;
        if (q < nchs) {
          const int e = 4 * q - ph;  // slab element index of the chunk's first float (head chunk: may be < 0)
          int rA = 0;
#pragma unroll
          for (int r = 1; r < GB; ++r) rA += (e >= r * C) ? 1 : 0;
          const int us = (rA + 1) * C - e;  // floats u >= us of the chunk belong to row rA+1
          const float2 tA = tab[rA], tB = tab[rA + 1];
          float4 x = v[c];
          x.x = soft1(x.x, us > 0 ? tA : tB);
          x.y = soft1(x.y, us > 1 ? tA : tB);
          x.z = soft1(x.z, us > 2 ? tA : tB);
          x.w = soft1(x.w, us > 3 ? tA : tB);
          slab4[q] = x;
        }
      }
    }
    __syncwarp();
    // gamma scatter: states that share a class are spread over rounds by their duplicate rank, so every round
    // is a conflict-free read-add-write on the row (repeated labels accumulate, quirk 6)
    const bool live = t < Tb;
    const RowGeom g = geom(t, tsl);
    float* xr = reinterpret_cast<float*>(g.srow) + g.off4;
    float gam[NSL];
    if (live) {
      const double* at = abt + seq * G::ABSEQ + ti * AS;
      const double* bt = at + TT * AS;
      const double s2 = s2v[seq];
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const int st = li + j * LPR;
        gam[j] = label(j) >= 0 ? (float)(at[st] * (bt[st] * s2)) : 0.f;
      }
    }
    const int nr = __reduce_max_sync(0xffffffffu, live ? max_rank : 0);
    for (int r = 0; r <= nr; ++r) {
      if (live) {
#pragma unroll
        for (int j = 0; j < NSL; ++j) {
          const int l = label(j);
          if (l >= 0 && (l >> kLabBits) == r) xr[l & kLabMask] += gam[j];
        }
      }
      if (r < nr) __syncwarp();
    }
  }
};

// ============================================================================ producer warp (TMA)
// lane i moves time step i of a tile
template <int TT>
struct Producer {
  const Problem& P;
  const StreamCfg& cfg;
  const Smem& S;
  const int lane;
  const int64_t b0;
  const uint32_t gbytes;  // bytes of the group's rows at one time step
  const uint64_t pol_keep;
  const uint64_t lim;     // one past the last logit

  __device__ __forceinline__ Producer(const Problem& P_, const StreamCfg& cfg_, const Smem& S_, int lane_, int64_t b0_, int gcnt)
      : P(P_), cfg(cfg_), S(S_), lane(lane_), b0(b0_), gbytes((uint32_t)gcnt * (uint32_t)P_.C * 4u),
        pol_keep(policy_evict_last()), lim(reinterpret_cast<uint64_t>(P_.logits) + (uint64_t)P_.T * P_.B * P_.C * 4u) {}

  __device__ __forceinline__ uint64_t elem_off(int t) const { return (((uint64_t)t * P.B + b0) * P.C) * 4u; }

  // rows of tile k, time steps [k*TT, k*TT+nvl) -> ring slot; 16-byte aligned superset of each time step's rows
  __device__ __forceinline__ void load_item(int slot, int k, int nvl, bool keep) const {
    uint64_t* bar = &S.sfull[slot];
    if (lane < nvl) {
      unsigned char* dst = S.ring + ((size_t)slot * TT + lane) * cfg.RSg;
      const uint64_t a = reinterpret_cast<uint64_t>(P.logits) + elem_off(k * TT + lane);
      const uint64_t a0 = a & ~uint64_t(15);
      uint64_t a1 = (a + gbytes + 15) & ~uint64_t(15);
      if (a1 > lim) {
        // the tensor's last rows end inside a 16-byte chunk: the bulk copy stops before it, the rest goes by hand
        a1 = lim & ~uint64_t(15);
        const float* src = reinterpret_cast<const float*>(a1);
        float* d = reinterpret_cast<float*>(dst + (a1 - a0));
        const int n = (int)((a + gbytes - a1) >> 2);
        for (int c = 0; c < n; ++c) d[c] = __ldg(src + c);
      }
      if (a1 > a0) {
        const uint32_t nb = (uint32_t)(a1 - a0);
        mbar_expect_tx(bar, nb);
        if (keep) bulk_g2s_hint(smem_u32(dst), a0, nb, smem_u32(bar), pol_keep);
        else bulk_g2s(smem_u32(dst), a0, nb, smem_u32(bar));
      }
    }
    __syncwarp();  // every lane's expect_tx (and hand-copied tail) precedes the one arrival that can end the phase
    if (lane == 0) mbar_arrive(bar);
  }

  // gradient rows of tile k, time steps [k*TT, k*TT+nvs): ring slot -> grad; the 16-byte aligned interior of
  // each time step goes out as one bulk store, at most 3 floats on either side by hand.  Every lane commits one
  // (possibly empty) bulk group per call, so that wait_group counts line up across lanes.
  __device__ __forceinline__ void store_item(int slot, int k, int nvs) const {
    if (lane < nvs) {
      const uint64_t g = reinterpret_cast<uint64_t>(P.grad) + elem_off(k * TT + lane);
      const uint64_t gend = g + gbytes;
      uint64_t g0 = (g + 15) & ~uint64_t(15), g1 = gend & ~uint64_t(15);
      const unsigned char* src = S.ring + ((size_t)slot * TT + lane) * cfg.RSg + (g & 15);  // image of byte g
      if (g1 > g0) {
        bulk_s2g(g0, smem_u32(src + (g0 - g)), (uint32_t)(g1 - g0));
      } else {
        g0 = gend; g1 = gend;  // everything by hand
      }
      for (uint64_t q = g; q < g0; q += 4) *reinterpret_cast<float*>(q) = *reinterpret_cast<const float*>(src + (q - g));
      for (uint64_t q = g1; q < gend; q += 4) *reinterpret_cast<float*>(q) = *reinterpret_cast<const float*>(src + (q - g));
    }
    bulk_commit();
  }
};

// ============================================================================ kernel
template <int NS, int LPR, int CPL, int MINB>
__global__ void __launch_bounds__(Geo<NS, LPR>::NTHREADS, MINB) nbctc_stream_kernel(const Problem P, const StreamCfg cfg) {
  using G = Geo<NS, LPR>;
  constexpr int TT = G::TT, Lpad = G::Lpad, PS = G::PS, AS = G::AS, GB = G::GB, NRW = G::NRW, NSLOT = kNSlot;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem S;
  S.sfull = reinterpret_cast<uint64_t*>(smem_raw + cfg.o_bar);
  S.info = reinterpret_cast<int*>(smem_raw + cfg.o_info);
  S.lab = reinterpret_cast<int*>(smem_raw + cfg.o_lab);
  S.lse = reinterpret_cast<float*>(smem_raw + cfg.o_lse);
  S.ckpt = reinterpret_cast<double*>(smem_raw + cfg.o_ckpt);
  S.cke = reinterpret_cast<int*>(smem_raw + cfg.o_cke);
  S.ptile = reinterpret_cast<float*>(smem_raw + cfg.o_ptile);
  S.ab = reinterpret_cast<double*>(smem_raw + cfg.o_ab);
  S.s2 = reinterpret_cast<double*>(smem_raw + cfg.o_s2);
  S.tab = reinterpret_cast<float2*>(smem_raw + cfg.o_tab);
  S.ring = smem_raw + cfg.o_ring;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b0 = (int64_t)blockIdx.x * GB;
  const int gcnt = (int)min((int64_t)GB, P.B - b0);

  // ---- per-sequence lengths, labels, validity (include/nbctc.h parity domain)
  if (tid < 2 * kMaxGB) S.info[2 * kMaxGB + tid] = 0;
  if (tid < NSLOT) mbar_init(&S.sfull[tid], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  for (int idx = tid; idx < GB * Lpad; idx += G::NTHREADS) {
    const int r = idx / Lpad, s = idx - r * Lpad;
    int l = 0;
    if (r < gcnt) {
      const int64_t Tb64 = P.in_len[b0 + r], Lb64 = P.tgt_len[b0 + r];
      if (seq_feasible(Tb64, Lb64, P.T, P.Lmax) && s < Lb64) {
        l = P.labels[(b0 + r) * P.Lmax + s];
        if (l < 0 || l >= P.C) { atomicOr(&S.info[2 * kMaxGB + r], 1); l = 0; }
      }
    }
    S.lab[idx] = l;
  }
  __syncthreads();
  if (tid < GB) {
    int Tb = 0, Lb = 0;
    if (tid < gcnt) {
      const int64_t Tb64 = P.in_len[b0 + tid], Lb64 = P.tgt_len[b0 + tid];
      if (seq_feasible(Tb64, Lb64, P.T, P.Lmax) && S.info[2 * kMaxGB + tid] == 0) {
        Tb = (int)Tb64; Lb = (int)Lb64;
      } else {
        P.loss[b0 + tid] = INFINITY;
      }
    }
    S.info[tid] = Tb;
    S.info[kMaxGB + tid] = Lb;
  }
  __syncthreads();
  // duplicate ranks (the loop reads the class bits of earlier states while later ones may already be packed)
  for (int idx = tid; idx < GB * Lpad; idx += G::NTHREADS) {
    const int r = idx / Lpad, s = idx - r * Lpad;
    int rank = 0;
    if (s < S.info[kMaxGB + r]) {
      const int l = S.lab[idx] & kLabMask;
      for (int q = 0; q < s; ++q) rank += ((S.lab[r * Lpad + q] & kLabMask) == l) ? 1 : 0;
      if (rank > 0) {
        atomicMax(&S.info[3 * kMaxGB + r], rank);
        S.lab[idx] = l | (rank << kLabBits);
      }
    }
  }
  __syncthreads();
  int Tg = 0;
#pragma unroll
  for (int r = 0; r < GB; ++r) Tg = max(Tg, S.info[r]);
  const int NTg = (Tg + TT - 1) / TT;
  const bool want_grad = P.grad != nullptr;
  const int total = want_grad ? 2 * NTg : NTg;
  const size_t slot_bytes = (size_t)TT * cfg.RSg;
  PROF_DECL

  // rows beyond the group's longest input: all-zero gradient, written directly (only ragged batches get here)
  if (want_grad && NTg * TT < P.T) {
    const int64_t n = (int64_t)gcnt * P.C;
    for (int64_t t = (int64_t)NTg * TT; t < P.T; ++t) {
      float* dst = P.grad + (t * P.B + b0) * P.C;
      for (int64_t c = tid; c < n; c += G::NTHREADS) dst[c] = 0.f;
    }
  }

#ifdef NBCTC_PROF
#define NBCTC_ITER_END()                                                                            \
  {                                                                                                 \
    const long long t_work_ = clock64();                                                            \
    PROF_SCOPE(6, __syncthreads(); prof_[5] += *reinterpret_cast<volatile int*>(S.info) & 0)         \
    if (cfg.prof != nullptr && blockIdx.x == gridDim.x / 2 && lane == 0 && it + 1 < 128) {           \
      cfg.prof[24 + ((it + 1) * 32 + warp) * 2] = t_work_ - prof_t0_;                                \
      cfg.prof[24 + ((it + 1) * 32 + warp) * 2 + 1] = clock64() - prof_t0_;                          \
    }                                                                                               \
  }
#else
#define NBCTC_ITER_END() __syncthreads();
#endif

  if (warp < GB) {
    // ======================================================================== chain warp of sequence `warp`
    const int seq = warp;
    const int Tb = S.info[seq], Lb = S.info[kMaxGB + seq];
    const int NTb = (Tb + TT - 1) / TT;
    const float wgt = (seq < gcnt) ? P.w_scalar * (P.seq_w ? P.seq_w[b0 + seq] : 1.f) : 0.f;
    double* ck = (cfg.ckpt_global ? cfg.ws_ckpt + ((size_t)min(b0 + seq, P.B - 1) * cfg.NTmax) * Lpad
                                  : S.ckpt + ((size_t)seq * cfg.NTmax) * Lpad) + (lane & 15);  // [NTmax][NS][16]
    int* cke = cfg.ckpt_global ? cfg.ws_cke + (size_t)min(b0 + seq, P.B - 1) * cfg.NTmax : S.cke + (size_t)seq * cfg.NTmax;
    ChainScal chain;
    double cx[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) cx[j] = 0.0;
    chain.carry = (lane == 0) ? 1.0 : 0.0;
    chain.zinv = 0.0;
    chain.Ea = 0; chain.Eb = 0; chain.Ez = 0;
    for (int it = -1; it <= total + 1; ++it) {
      if (it >= 0 && it < total) {
        const int buf = it & 1;
        const float* pt = S.ptile + (size_t)(buf * GB + seq) * G::PSEQ;
        if (it < NTg) {
          if (it < NTb) {
            PROF_SCOPE(0, chain_phase1<NS, TT, PS>(cx, chain, lane, Tb, ck, cke, it, pt))
            if (it == NTb - 1) chain_readout<NS>(cx, chain, lane, Lb, &P.loss[b0 + seq], wgt);
          }
        } else {
          const int k = 2 * NTg - 1 - it;
          if (k < NTb) {
            PROF_SCOPE(1, chain_phase2<NS, TT, PS, AS>(cx, chain, lane, Tb, ck, cke, k, pt,
                                                       S.ab + (size_t)(buf * GB + seq) * G::ABSEQ, &S.s2[buf * GB + seq]))
          }
        }
      }
      NBCTC_ITER_END()
    }
    PROF_DUMP(0)
  } else if (warp < GB + NRW) {
    // ======================================================================== row warp of time step `ti`
    const int ti = warp - GB;
    const Rows<NS, LPR, CPL> rows(P, cfg, S, lane, ti, gcnt, b0);
    for (int it = -1; it <= total + 1; ++it) {
      const int a = it + 1;  // ahead item
      if (a >= 0 && a < total) {
        const int slot = a % NSLOT;
        PROF_SCOPE(0, mbar_wait(&S.sfull[slot], (uint32_t)(a / NSLOT) & 1u))
        float* pt = S.ptile + (size_t)((a & 1) * GB) * G::PSEQ;
        unsigned char* tsl = S.ring + slot * slot_bytes + (size_t)ti * cfg.RSg;
        if (a < NTg) {
          PROF_SCOPE(1, rows.forward_step(a * TT + ti, tsl, pt))
        } else {
          PROF_SCOPE(2, rows.emit_step((2 * NTg - 1 - a) * TT + ti, tsl, pt))
        }
      }
      const int g = it - 1;  // behind item
      if (g >= NTg && g < total) {
        const int t = (2 * NTg - 1 - g) * TT + ti;
        if (t < P.T) {
          const int buf = g & 1;
          PROF_SCOPE(3, rows.grad_step(t, S.ring + (g % NSLOT) * slot_bytes + (size_t)ti * cfg.RSg,
                                       S.ab + (size_t)(buf * GB) * G::ABSEQ, S.s2 + buf * GB))
        }
        fence_proxy_async();  // the slot is read by the async proxy (bulk store) after the barrier
      }
      NBCTC_ITER_END()
    }
    PROF_DUMP(1)
  } else {
    // ======================================================================== producer warp
    const Producer<TT> prod(P, cfg, S, lane, b0, gcnt);
    int next_load = 0;
    for (int it = -1; it <= total + 1; ++it) {
      const int s = it - 2;  // item whose gradient rows are complete
      if (s >= NTg && s < total) {
        const int k = 2 * NTg - 1 - s;
        PROF_SCOPE(0, prod.store_item(s % NSLOT, k, min(TT, (int)P.T - k * TT)))
      }
      while (next_load < total && next_load <= it + NSLOT) {
        const int occ = next_load - NSLOT;  // previous occupant of the slot
        if (occ >= 0) {
          const int free_at = occ < NTg ? occ : occ + 3;
          if (free_at > it) break;
          if (occ >= NTg) PROF_SCOPE(1, bulk_wait_read<1>())  // all but the store committed just above have been read
        }
        const int k = next_load < NTg ? next_load : 2 * NTg - 1 - next_load;
        PROF_SCOPE(2, prod.load_item(next_load % NSLOT, k, min(TT, Tg - k * TT), next_load < NTg))
        ++next_load;
      }
      NBCTC_ITER_END()
    }
    bulk_wait_read<0>();  // the ring must outlive the last bulk store's reads
    PROF_DUMP(2)
  }
#undef NBCTC_ITER_END
}

template <int NS, int LPR, int CPL>
int launch_inst(const Problem& p, const StreamCfg& cfg, cudaStream_t stream) {
  using G = Geo<NS, LPR>;
  const unsigned groups = (unsigned)((p.B + G::GB - 1) / G::GB);
  if constexpr (G::GB <= 2) {
    if (cfg.ctas_per_sm >= 2) {
      auto kern = nbctc_stream_kernel<NS, LPR, CPL, 2>;
      if (cfg.smem_bytes > 48 * 1024)
        NBCTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes));
      NBCTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      kern<<<groups, G::NTHREADS, cfg.smem_bytes, stream>>>(p, cfg);
      NBCTC_LAUNCH_CHECK();
      return NBCTC_OK;
    }
  }
  {
    auto kern = nbctc_stream_kernel<NS, LPR, CPL, 1>;
    if (cfg.smem_bytes > 48 * 1024)
      NBCTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes));
    NBCTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    kern<<<groups, G::NTHREADS, cfg.smem_bytes, stream>>>(p, cfg);
  }
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

template <int NS>
int launch_ns(const Problem& p, const StreamCfg& cfg, cudaStream_t stream) {
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
  } else if (cfg.LPR == 16) {
    switch (cfg.CPL) {
      case 2: return launch_inst<NS, 16, 2>(p, cfg, stream);
      case 3: return launch_inst<NS, 16, 3>(p, cfg, stream);
      case 4: return launch_inst<NS, 16, 4>(p, cfg, stream);
    }
  } else if (cfg.LPR == 32) {
    switch (cfg.CPL) {
      case 3: return launch_inst<NS, 32, 3>(p, cfg, stream);
      case 4: return launch_inst<NS, 32, 4>(p, cfg, stream);
      case 6: return launch_inst<NS, 32, 6>(p, cfg, stream);
      case 8: return launch_inst<NS, 32, 8>(p, cfg, stream);
    }
  }
  set_error("no stream kernel instance for LPR=%d CPL=%d", cfg.LPR, cfg.CPL);
  return NBCTC_ERR_UNSUPPORTED;
}

}  // namespace stream
#endif  // __CUDACC__

}  // namespace nbctc
