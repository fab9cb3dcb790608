// Block-streaming no-blank CTC forward+backward kernel for sm_100a (overview in nbctc_stream.cu).
//
// One CTA owns a GROUP of GB = 32/LPR batch-adjacent sequences and walks their lattice in tiles of TT time steps.
// The rows (t, b0..b0+GB) of one time step are contiguous in the (T,B,C) tensor ("slab").  Row warp w owns time
// step w of every tile: it moves its slab with TMA bulk copies (global -> ring slot -> global) and no other warp
// ever touches it.  Chain warp r owns sequence r.  The two roles meet in small shared-memory tiles and advance
// in lock step, one __syncthreads() per tile, nobody polls.
//
//   phase 1 (tiles upwards)    row warps    slab of item it+1: row log-partition (NoBlankCTC.py:136), then the slab
//                                           becomes w*softmax(x) IN PLACE (the exponentials of the log-partition
//                                           are reused) and goes to the gradient tensor with one bulk store;
//                                           emissions p_t(s) = softmax(x_t)[label_s] (NoBlankCTC.py:96-102) are
//                                           gathered from the finished slab -> p-tile
//                              chain warps  item it: alpha over the tile (NoBlankCTC.py:71-87), 16 lanes x NS
//                                           states in float64, linear domain, exact power-of-two rescaling and one
//                                           checkpoint per tile
//   phase 2 (tiles downwards)  row warps    the finished slabs come back from the gradient tensor (L2 hits).  Item it+1:
//                                           the same emissions again (bit-identical to phase 1); item it-1:
//                                           gamma = alpha*beta/Z from the chain's tiles, -w*gamma added to the slab
//                                           in shared memory (repeated labels accumulate), and
//                                           the slab goes back with one bulk store
//                              chain warps  item it: lanes 0-15 replay alpha inside the tile from the checkpoint
//                                           while lanes 16-31 run beta in the same instructions (beta is kept in
//                                           reversed state order)
//
// HBM traffic: the logits are read once and the gradient is written once, as long as the gradient rows of the
// sequences in flight stay in L2 between their two phases.
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
  int ckpt_global;
  int ctas_per_sm;
  uint32_t o_bar, o_info, o_lab, o_ckpt, o_cke, o_ptile, o_s2, o_pub, o_ab, o_ring, smem_bytes;
  double* ws_ckpt;  // [B][NTmax][Lpad]
  int* ws_cke;      // [B][NTmax][W]
  int* floor_flag;  // [B] or null: set when an emission of a live state fell below the float32 floor (nbctc_logdom.cu)
  long long* prof;  // role profiler output (NBCTC_PROF builds), else null: [3][8] buckets, then [160][32][2] trace
};

template <int NS, int LPR>
struct Geo {
  // time steps per tile.  Lpad = 256: two steps only, so that the tiles are small and three CTAs (one sequence
  // each) share an SM -- these chains are long and the only parallelism is across sequences
  static constexpr int TT = NS >= 16 ? 2 : 8;
  static constexpr int Lpad = 16 * NS;
  static constexpr int GB = 32 / LPR;          // sequences per CTA
  static constexpr int NRW = TT;               // row warps: one per time step of a tile
  static constexpr int NSL = Lpad / LPR;       // states per lane in the emission gather / gamma scatter
  static constexpr int PS = Lpad + 8;          // p-tile row stride (floats)
  static constexpr int AS = Lpad + 8;          // alpha/beta tile row stride (doubles)
  static constexpr int PSEQ = TT * PS + 8;     // p-tile floats per sequence (+8: lane groups hit distinct banks)
  static constexpr int ABSEQ = 2 * TT * AS + 8;  // alpha+beta tile doubles per sequence
  static constexpr int W = NS > 2 ? 32 : 16;   // chain lanes per direction (NS = Lpad/16 names the instance)
  static constexpr int CNS = Lpad / W;         // chain states per lane
  static constexpr int NCW = W == 32 ? 2 : 1;  // chain warps per sequence: W = 32 -> one for alpha, one for beta
  static constexpr int NCHAIN = GB * NCW;
  static constexpr int NMW = 2;              // mover warps: TMA issue for TT/NMW time steps each
  static constexpr int NTHREADS = 32 * (NCHAIN + NRW + NMW);
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
__device__ __forceinline__ void bulk_s2g_hint(uint64_t dst_gmem, uint32_t src_smem, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
               "r"(src_smem), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
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

// what the alpha warp of a sequence publishes for its beta warp (two chain warps per sequence, W = 32)
struct ChainPub {
  double zinv;
  int Ez, Eb;
};

struct Smem {
  uint64_t* sfull;  // [kNSlot][TT] "this time step's slab has landed" barriers
  int* info;        // [0..GB) T_b, [kMaxGB..) L_b, [2*kMaxGB..) bad-label flags, [3*kMaxGB..) largest duplicate rank
  int* lab;         // [GB][Lpad]
  double* ckpt;     // [GB][NTmax][Lpad]   (or in the workspace)
  int* cke;         // [GB][NTmax][W] lane scales of the checkpoints
  float* ptile;     // [2][GB][PSEQ]
  double* s2;       // [2][GB]
  ChainPub* pub;    // [GB] alpha warp -> beta warp hand-over (W = 32)
  double* ab;       // [2][GB][ABSEQ]
  unsigned char* ring;  // [kNSlot][TT][RSg]
};

// ============================================================================ chain warps
// W = 16: one warp per sequence; lanes 0-15 run alpha, lanes 16-31 run beta (phase 2) in the same instructions.
// W = 32 (Lmax > 32): two warps per sequence, 32 lanes each: one runs alpha (phase 1 + replay), the other beta.
// In both layouts position q of a direction's W lanes x NS states is state q for alpha and state Lpad-1-q for beta,
// so both shift the same way.

// ---- lane scales.  The states of a direction are spread over W lanes x NS states; they can differ by thousands of
// binary orders of magnitude across the lattice (a model that is sure of one label for the whole sequence), far more
// than the float64 range.  Every lane therefore carries its own exact power-of-two scale e_l (state = x * 2^e_l),
// renewed once per tile, and the neighbour's state enters a lane through fac = 2^(e_{l-1} - e_l).
constexpr int kSent = -(1 << 28);  // "no exponent": an all-zero lane
// exact 2^e; 0 below the normal range, 2^1023 above
__device__ __forceinline__ double pow2z(int e) {
  if (e < -1022) return 0.0;
  return __hiloint2double((1023 + min(e, 1023)) << 20, 0);
}
// exponent field - 1023 of the largest of NS non-negative values; kSent if that is zero / denormal / not finite
template <int NS>
__device__ __forceinline__ int top_exponent(const double (&v)[NS]) {
  int hi = __double2hiint(v[0]);
#pragma unroll
  for (int j = 1; j < NS; ++j) hi = max(hi, __double2hiint(v[j]));
  const int ef = hi >> 20;
  return (ef > 0 && ef < 0x7ff) ? ef - 1023 : kSent;
}
// Largest scale step between neighbouring lanes for tiles of TT steps: mass crosses at most ceil(TT/NS) lanes between
// two rescales and gains 2^DEC of scaled magnitude per crossing at worst, which must stay inside the float64 range.
template <int NS, int TT>
struct LaneDec {
  static constexpr int value = (TT + NS - 1) / NS >= 4 ? 208 : (TT + NS - 1) / NS >= 2 ? 420 : 850;
};
// New lane scale e_l = max_{k<=l}(A_k - DEC (l-k)), A_k = absolute exponent of lane k's largest state: every lane is
// scaled to its own magnitude unless mass from a much larger lane upstream is about to arrive (mass only moves to
// higher positions).  Values that fall 2^-1022 below the lane scale flush to zero; they are below float64 resolution
// of what that mass turns them into.
template <int NS, int W, int DEC>
__device__ __forceinline__ void lane_rescale(double (&x)[NS], int& e, double& fac, int hl) {
  const int te = top_exponent<NS>(x);
  int env = te == kSent ? kSent : e + te;
#pragma unroll
  for (int o = 1; o < W; o <<= 1) {
    const int sh = __shfl_up_sync(0xffffffffu, env, o, W);
    if (hl >= o && sh > kSent / 2) env = max(env, sh - DEC * o);
  }
  const int en = env > kSent / 2 ? env : e;
  const double sc = pow2z(e - en);
#pragma unroll
  for (int j = 0; j < NS; ++j) x[j] *= sc;
  e = en;
  const int eu = __shfl_up_sync(0xffffffffu, en, 1, W);
  fac = hl == 0 ? 0.0 : pow2z(eu - en);
}

// emissions of the lane's NS states for one row of a p-tile, in the lane's own state order (rev: beta)
template <int NS, int W>
__device__ __forceinline__ void load_p(const float* row, int hl, bool rev, double (&p)[NS]) {
  const float* src = row + (rev ? (W - 1 - hl) * NS : hl * NS);
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

// x(s) <- (x(s) + x(s-1)) * p(s) in the lane's (possibly reversed) state order, written as
// x(s) <- fma(x(s-1), p(s), x(s)*p(s)): the products x(s)*p(s) do not wait for the neighbour lane's state, so the
// dependent path of a step is one shuffle + one DFMA for the lane's first state and one DFMA for the others
// (a shuffle, a DADD and a DMUL in the textbook form).  alpha: x = alpha (NoBlankCTC.py:73-85).  beta:
// x(s) = beta_t(s) p_t(s) and, with kSum, sum(s) = x(s) + x(s-1) = beta_t(s) (off the dependent path).
// fac = 2^(e_neighbour - e_lane) brings the neighbour lane's last state into this lane's scale (0 for the first lane
// of a direction, whose shuffle returns its own finite value).
// kFirst: first step of a tile -- `carry` enters position 0 of the direction, in the lane's own scale (the virtual
// start state: NoBlankCTC.py:92-93 and the t>0 guard at :75).
template <int NS, int W, bool kSum, bool kFirst>
__device__ __forceinline__ void chain_step(double (&x)[NS], double (&sum)[NS], const double (&p)[NS], int hl, double carry, double fac) {
  double t[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) t[j] = x[j] * p[j];
  double up = __shfl_up_sync(0xffffffffu, x[NS - 1], 1, W);
  double f0 = fac;
  if (kFirst) {
    if (hl == 0) {
      up = carry;
      f0 = 1.0;
    }
  }
  const double p0 = p[0] * f0;
  if (kSum) {
#pragma unroll
    for (int j = NS - 1; j >= 1; --j) sum[j] = x[j] + x[j - 1];
    sum[0] = fma(up, f0, x[0]);
  }
#pragma unroll
  for (int j = NS - 1; j >= 1; --j) x[j] = fma(x[j - 1], p[j], t[j]);
  x[0] = fma(up, p0, t[0]);
}
// run-time `first` (generic loops)
template <int NS, int W, bool kSum>
__device__ __forceinline__ void chain_step_rt(double (&x)[NS], double (&sum)[NS], const double (&p)[NS], int hl, double carry,
                                              double fac, bool first) {
  if (first) chain_step<NS, W, kSum, true>(x, sum, p, hl, carry, fac);
  else chain_step<NS, W, kSum, false>(x, sum, p, hl, 0.0, fac);
}

// Chain-warp state lives in plain registers of the kernel body (passed by reference to force-inlined functions).
struct ChainScal {
  double carry, zinv, fac;
  int e, Ez;  // this lane's scale; exponent of Z
};

// ---- phase 1, tile k: alpha over the tile's steps (W = 16: lanes 16-31 carry zeros).  Checkpoint of tile k: the
// states ck[(k*NS+j)*W] and the lane scales cke[k*W + lane].
template <int NS, int W, int TT, int PS>
__device__ __forceinline__ void chain_phase1(double (&x)[NS], ChainScal& c, int lane, int Tb, double* ck, int* cke, int k,
                                             const float* __restrict__ pt) {
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
  const int nv = min(TT, Tb - k * TT);
  if (NS * TT <= 16 && nv == TT) {
    // whole tile of emissions in registers ahead of the dependent loop
    double pr[TT][NS];
#pragma unroll
    for (int i = 0; i < TT; ++i) load_p<NS, W>(pt + i * PS, hl, false, pr[i]);
#pragma unroll
    for (int i = 0; i < TT; ++i) {
      if (i == 0) chain_step<NS, W, false, true>(x, sum, pr[i], hl, c.carry, c.fac);
      else chain_step<NS, W, false, false>(x, sum, pr[i], hl, 0.0, c.fac);
    }
  } else {
#pragma unroll 2
    for (int i = 0; i < nv; ++i) {
      double pf[NS];
      load_p<NS, W>(pt + i * PS, hl, false, pf);
      chain_step_rt<NS, W, false>(x, sum, pf, hl, c.carry, c.fac, i == 0);
    }
  }
  c.carry = 0.0;
}

// beta start state: position q holds state Lpad-1-q; x = u_t(s) = beta_t(s) p_t(s).  Virtual start
// u_{T_b}(L_b) = 1 gives beta_{T_b-1}(L_b-1) = 1 without a branch (state L_b has p = 0 and alpha = 0).
template <int NS, int W>
__device__ __forceinline__ void chain_beta_init(double (&x)[NS], ChainScal& c, int hl, int Lb) {
  constexpr int Lpad = W * NS;
  c.e = 0;
  c.fac = hl == 0 ? 0.0 : 1.0;
#pragma unroll
  for (int j = 0; j < NS; ++j) x[j] = (Lpad - 1 - (hl * NS + j) == Lb) ? 1.0 : 0.0;
  c.carry = (hl == 0 && Lb == Lpad) ? 1.0 : 0.0;
}

// ---- read-out after the sequence's last phase-1 tile (NoBlankCTC.py:58-68,:139).  W = 16: the beta half of the
// warp starts here; W = 32: the alpha warp publishes 1/Z and its exponent for the beta warp.
template <int NS, int W>
__device__ __forceinline__ void chain_readout(double (&x)[NS], ChainScal& c, int lane, int Lb, float* loss_out, float wgt,
                                              ChainPub* pub) {
  const int hl = lane & (W - 1);
  const int sl = Lb - 1;
  // x[sl % NS] without a dynamic index (which would push the state array into local memory)
  const int rj = sl % NS;
  double mine = x[0];
#pragma unroll
  for (int j = 1; j < NS; ++j) mine = (rj >= j) ? x[j] : mine;
  double zhat = __shfl_sync(0xffffffffu, mine, sl / NS);
  c.Ez = __shfl_sync(0xffffffffu, c.e, sl / NS);
  // normalise to [1, 2): the lane's states may have shrunk since its last rescale
  const int ezf = __double2hiint(zhat) >> 20;
  const bool ok = zhat > 0.0 && ezf > 0 && ezf < 0x7ff;
  if (ok) {
    zhat *= pow2z(1023 - ezf);
    c.Ez += ezf - 1023;
  }
  if (lane == 0) *loss_out = ok ? (float)(-(log(zhat) + (double)c.Ez * 0.6931471805599453)) : INFINITY;
  c.zinv = ok ? (double)wgt / zhat : 0.0;  // sequence weight folded into gamma
  if (W == 16) {
    if (lane >= 16) chain_beta_init<NS, W>(x, c, hl, Lb);
  } else if (lane == 0) {
    pub->zinv = c.zinv;
    pub->Ez = c.Ez;
    pub->Eb = 0;
  }
}

// ---- phase 2, tile k: beta (isb) / alpha replay (!isb) from the checkpoint (states ckv; eck = the checkpoint's
// scale of this lane for the alpha direction, of the alpha lane that holds the same states for the beta direction;
// unused for k = 0).  ab tile: alpha_t(s) in the alpha lane's scale, and beta_t(s) 2^(e_beta + e_alpha - Ez) w / Zhat,
// so that gamma_t(s) = alpha-entry * beta-entry for the row warps.  (The beta entry is kept finite: where the alpha
// entry has flushed to zero -- a state far below the mass arriving from upstream -- gamma is 0, not 0 * inf.)
template <int NS, int W, int TT, int PS, int AS>
__device__ __forceinline__ void chain_phase2(double (&x)[NS], ChainScal& c, int lane, bool isb, int Tb, const double (&ckv)[NS],
                                             int eck, int k, const float* __restrict__ pt, double* __restrict__ abt) {
  constexpr int Lpad = W * NS;
  const int hl = lane & (W - 1);
  double sum[NS];
  if (!isb) {
    // alpha replay: from the virtual start state (k = 0) or the checkpoint
    c.e = k == 0 ? 0 : eck;
    c.carry = (k == 0 && hl == 0) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < NS; ++j) x[j] = k == 0 ? 0.0 : ckv[j];
  }
  {
    // (every lane takes part in the shuffle; the beta direction keeps its own fac)
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
  const int nv = min(TT, Tb - k * TT);
  // position -> state index of this lane's slots
  const int s0 = isb ? (Lpad - 1 - hl * NS) : hl * NS;
  const int sdir = isb ? -1 : 1;
  double* dst = abt + (isb ? TT * AS : 0) + s0;  // alpha tile, then beta tile
  if (NS * TT <= 16 && nv == TT) {
    double pr[TT][NS];
#pragma unroll
    for (int jj = 0; jj < TT; ++jj) load_p<NS, W>(pt + (isb ? TT - 1 - jj : jj) * PS, hl, isb, pr[jj]);
#pragma unroll
    for (int jj = 0; jj < TT; ++jj) {
      const int i = isb ? (TT - 1 - jj) : jj;  // alpha walks up the tile, beta walks down
      if (jj == 0) chain_step<NS, W, true, true>(x, sum, pr[jj], hl, c.carry, c.fac);
      else chain_step<NS, W, true, false>(x, sum, pr[jj], hl, 0.0, c.fac);
#pragma unroll
      for (int j = 0; j < NS; ++j) dst[i * AS + sdir * j] = isb ? fmin(sum[j] * bs, 1e300) * bs2 : x[j];
    }
  } else {
#pragma unroll 2
    for (int jj = 0; jj < nv; ++jj) {
      const int i = isb ? (nv - 1 - jj) : jj;
      double pf[NS];
      load_p<NS, W>(pt + i * PS, hl, isb, pf);
      chain_step_rt<NS, W, true>(x, sum, pf, hl, c.carry, c.fac, jj == 0);
#pragma unroll
      for (int j = 0; j < NS; ++j) dst[i * AS + sdir * j] = isb ? fmin(sum[j] * bs, 1e300) * bs2 : x[j];
    }
  }
  c.carry = 0.0;
  // every lane takes part in the group-wide shuffles; only the beta direction keeps the result
  int e2 = c.e;
  double f2 = c.fac;
  lane_rescale<NS, W, LaneDec<NS, TT>::value>(x, e2, f2, hl);
  if (isb) {
    c.e = e2;
    c.fac = f2;
  }
}

// ============================================================================ row warps
// Geometry of one (t,b) row seen as 16-byte chunks of its slab: the row starts `off4` floats into chunk 0 and ends
// `rem` floats into chunk nch-1 (rows are only 4-byte aligned when C % 4 != 0, and neighbouring sequences of the
// group share their boundary chunks).
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
// floats [lo, hi) of a chunk (a boundary chunk belongs to two rows, i.e. two lane groups)
__device__ __forceinline__ void store_part(float4* dst, const float4& y, int lo, int hi) {
  float* e = reinterpret_cast<float*>(dst);
  if (lo <= 0 && hi > 0) e[0] = y.x;
  if (lo <= 1 && hi > 1) e[1] = y.y;
  if (lo <= 2 && hi > 2) e[2] = y.z;
  if (lo <= 3 && hi > 3) e[3] = y.w;
}
__device__ __forceinline__ float4 scale4(const float4& v, float s) { return make_float4(v.x * s, v.y * s, v.z * s, v.w * s); }

// One row warp = one time step of every tile; lane group (LPR lanes) = one sequence of the CTA's group.
template <int NS, int LPR, int CPL>
struct Rows {
  using G = Geo<NS, LPR>;
  static constexpr int TT = G::TT, Lpad = G::Lpad, PS = G::PS, AS = G::AS, GB = G::GB, NSL = G::NSL;
  static constexpr bool kLabRegs = NSL <= 8;
  static constexpr int SEG = LPR * CPL;  // chunks per row segment

  const Problem& P;
  const StreamCfg& cfg;
  const Smem& S;
  const int lane, li, seq;  // seq = lane group = sequence of the group
  const int ti;             // this warp's time step inside a tile
  const int gcnt;
  const int64_t b0;
  const int Tb, Lb, C, max_rank;
  const uint32_t gbytes;    // bytes of the group's rows at one time step
  const float wgt;          // gradient weight of the sequence; weff = wgt, or 1 where wgt == 0 (see kernel tail)
  const float weff, winv;
  const int* lab_seq;
  const int ph_fixed;       // slab phase if it does not depend on t, else -1
  int labr[kLabRegs ? NSL : 1];  // this lane's labels (-1 = no state)

  __device__ __forceinline__ Rows(const Problem& P_, const StreamCfg& cfg_, const Smem& S_, int lane_, int ti_, int gcnt_,
                                  int64_t b0_, float wgt_)
      : P(P_), cfg(cfg_), S(S_), lane(lane_), li(lane_ & (LPR - 1)), seq(lane_ / LPR), ti(ti_), gcnt(gcnt_), b0(b0_),
        Tb(S_.info[lane_ / LPR]), Lb(S_.info[kMaxGB + lane_ / LPR]), C((int)P_.C), max_rank(S_.info[3 * kMaxGB + lane_ / LPR]),
        gbytes((uint32_t)gcnt_ * (uint32_t)P_.C * 4u), wgt(wgt_), weff(wgt_ != 0.f ? wgt_ : 1.f),
        winv(1.f / (wgt_ != 0.f ? wgt_ : 1.f)), lab_seq(S_.lab + (lane_ / LPR) * Lpad),
        ph_fixed((((unsigned)P_.B * (unsigned)P_.C) & 3u) == 0 ? (int)((((unsigned)b0_ & 3u) * ((unsigned)P_.C & 3u)) & 3u) : -1) {
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

  __device__ __forceinline__ uint64_t elem_off(int t) const { return (((uint64_t)t * P.B + b0) * P.C) * 4u; }
  // float index (0..3) of the slab's first element inside its first 16-byte chunk (both tensors are 16-byte aligned)
  __device__ __forceinline__ int slab_phase(int t) const {
    if (ph_fixed >= 0) return ph_fixed;  // B*C % 4 == 0: the same phase at every time step
    return (int)(((((unsigned)t & 3u) * ((unsigned)P.B & 3u) + ((unsigned)b0 & 3u)) * ((unsigned)C & 3u)) & 3u);
  }
  __device__ __forceinline__ unsigned char* slab(int slot) const { return S.ring + ((size_t)slot * TT + ti) * cfg.RSg; }
  // this lane group's row at time t inside the slab `tsl`
  __device__ __forceinline__ RowGeom geom(int t, unsigned char* tsl) const {
    const int fidx = slab_phase(t) + seq * C;  // float index of the row inside the slab's chunks
    RowGeom g;
    g.off4 = fidx & 3;
    g.nch = (g.off4 + C + 3) >> 2;
    g.rem = g.off4 + C - 4 * (g.nch - 1);
    g.srow = reinterpret_cast<float4*>(tsl) + (fidx >> 2);
    return g;
  }

  // ---------------------------------------------------------------- TMA (lane 0 only)
  // rows of time step t of `base` (logits, or the gradient in phase 2) -> ring slot: 16-byte aligned superset
  __device__ __forceinline__ void issue_load(int slot, int t, const float* base, uint64_t pol) const {
    uint64_t* bar = &S.sfull[slot * TT + ti];
    unsigned char* dst = slab(slot);
    const uint64_t a = reinterpret_cast<uint64_t>(base) + elem_off(t);
    const uint64_t lim = reinterpret_cast<uint64_t>(base) + (uint64_t)P.T * P.B * P.C * 4u;
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
      bulk_g2s_hint(smem_u32(dst), a0, (uint32_t)(a1 - a0), smem_u32(bar), pol);
      mbar_arrive_expect_tx(bar, (uint32_t)(a1 - a0));  // the phase cannot complete before this arrival
    } else {
      mbar_arrive(bar);
    }
  }
  // finished slab of time step t -> gradient rows: the 16-byte aligned interior as one bulk store, at most 3
  // floats on either side by hand
  __device__ __forceinline__ void issue_store(int slot, int t, uint64_t pol) const {
    const uint64_t g = reinterpret_cast<uint64_t>(P.grad) + elem_off(t);
    const uint64_t gend = g + gbytes;
    uint64_t g0 = (g + 15) & ~uint64_t(15), g1 = gend & ~uint64_t(15);
    const unsigned char* src = slab(slot) + (g & 15);  // shared-memory image of byte g
    if (g1 > g0) {
      bulk_s2g_hint(g0, smem_u32(src + (g0 - g)), (uint32_t)(g1 - g0), pol);
    } else {
      g0 = gend; g1 = gend;  // everything by hand
    }
    for (uint64_t q = g; q < g0; q += 4) *reinterpret_cast<float*>(q) = *reinterpret_cast<const float*>(src + (q - g));
    for (uint64_t q = g1; q < gend; q += 4) *reinterpret_cast<float*>(q) = *reinterpret_cast<const float*>(src + (q - g));
    bulk_commit();
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

  // Chunk slot c of a single-segment row is chunk q = li + c*LPR.  CPL <= ceil(nch_max/LPR) + 1 and nch varies by
  // at most one between the rows of a group, so the slots c <= CPL-3 always hold a full chunk of the row that is
  // not its last one: no bounds check, no tail handling (the head chunk is slot 0 of lane 0).
  static __device__ __forceinline__ constexpr bool slot_is_inner(int c) { return c + 3 <= CPL; }
  __device__ __forceinline__ float4 load_slot(const RowGeom& g, int c) const {
    const int q = li + c * LPR;
    if (slot_is_inner(c)) {
      float4 v = g.srow[q];
      if (c == 0 && li == 0) mask_head(v, g.off4);
      return v;
    }
    return load_chunk(g, q);
  }
  __device__ __forceinline__ void store_slot(const RowGeom& g, int c, const float4& y) const {
    const int q = li + c * LPR;
    if (slot_is_inner(c) && c > 0) g.srow[q] = y;
    else store_chunk(g, q, y);
  }
  // chunk q of the row -> registers, floats of neighbouring rows (and chunks beyond the row) as -inf
  __device__ __forceinline__ float4 load_chunk(const RowGeom& g, int q) const {
    float4 v = make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
    if (q < g.nch) {
      v = g.srow[q];
      if (q == 0) mask_head(v, g.off4);
      if (q == g.nch - 1) mask_tail(v, g.rem);
    }
    return v;
  }
  // chunk q of the row <- y; only this row's floats of a boundary chunk
  __device__ __forceinline__ void store_chunk(const RowGeom& g, int q, const float4& y) const {
    if (q < g.nch) {
      const int lo = q == 0 ? g.off4 : 0, hi = q == g.nch - 1 ? g.rem : 4;
      if (lo == 0 && hi == 4) g.srow[q] = y;
      else store_part(g.srow + q, y, lo, hi);
    }
  }

  // emissions p_t(s) = softmax(x_t)[label_s] = y[label_s] / w from the finished row (shared memory in phase 1,
  // the gradient tensor in phase 2: the same float either way) -> p-tile row
  __device__ __forceinline__ float emission(float y, int l) const { return l >= 0 ? fmaxf(y * winv, kPMin) : 0.f; }

  // ---------------------------------------------------------------- phase 1: one slab
  // row log-partition, slab -> w*softmax in place (zeros beyond input_length), emissions
  __device__ __forceinline__ void forward_step(int t, unsigned char* tsl, float* pt) const {
    const bool act = t < Tb;  // Tb = 0 for sequences outside the group / the parity domain
    const RowGeom g = geom(t, tsl);
    if (cfg.NSEG == 1) {
      float4 v[CPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c) v[c] = load_slot(g, c);
      float m_l = kNegInf;
#pragma unroll
      for (int c = 0; c < CPL; ++c) m_l = fmaxf(m_l, fmaxf(fmaxf(v[c].x, v[c].y), fmaxf(v[c].z, v[c].w)));
      const float mb = (m_l > kNegInf) ? m_l * kLog2e : 0.f;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        v[c].x = ex2f(fmaf(v[c].x, kLog2e, -mb)); s0 += v[c].x;
        v[c].y = ex2f(fmaf(v[c].y, kLog2e, -mb)); s1 += v[c].y;
        v[c].z = ex2f(fmaf(v[c].z, kLog2e, -mb)); s2 += v[c].z;
        v[c].w = ex2f(fmaf(v[c].w, kLog2e, -mb)); s3 += v[c].w;
      }
      const float m = group_max(m_l);
      const float cf = (m_l > kNegInf) ? ex2f((m_l - m) * kLog2e) : 0.f;  // this lane's exponentials -> row maximum
      const float s = group_sum(((s0 + s1) + (s2 + s3)) * cf);
      const float sc = act ? __fdividef(weff * cf, s) : 0.f;  // 1 <= s <= C: the fast division is exact to 2 ulp
#pragma unroll
      for (int c = 0; c < CPL; ++c) store_slot(g, c, scale4(v[c], sc));
    } else {
      // long rows: online max/sum over the segments, then a second pass over the slab
      float m_l = kNegInf, s_l = 0.f;
      for (int seg = 0; seg < cfg.NSEG; ++seg) {
        float4 v[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) v[c] = load_chunk(g, seg * SEG + li + c * LPR);
        float mm = m_l;
#pragma unroll
        for (int c = 0; c < CPL; ++c) mm = fmaxf(mm, fmaxf(fmaxf(v[c].x, v[c].y), fmaxf(v[c].z, v[c].w)));
        if (mm > kNegInf) {
          const float mb = mm * kLog2e;
          float s = 0.f;
#pragma unroll
          for (int c = 0; c < CPL; ++c)
            s += (ex2f(fmaf(v[c].x, kLog2e, -mb)) + ex2f(fmaf(v[c].y, kLog2e, -mb))) +
                 (ex2f(fmaf(v[c].z, kLog2e, -mb)) + ex2f(fmaf(v[c].w, kLog2e, -mb)));
          s_l = (m_l > kNegInf ? s_l * ex2f((m_l - mm) * kLog2e) : 0.f) + s;
          m_l = mm;
        }
      }
      const float m = group_max(m_l);
      const float s = group_sum(m_l > kNegInf ? s_l * ex2f((m_l - m) * kLog2e) : 0.f);
      const float mb = m * kLog2e;
      const float sc = act ? weff / s : 0.f;
      for (int seg = 0; seg < cfg.NSEG; ++seg) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const int q = seg * SEG + li + c * LPR;
          float4 v = load_chunk(g, q);
          v.x = ex2f(fmaf(v.x, kLog2e, -mb)); v.y = ex2f(fmaf(v.y, kLog2e, -mb));
          v.z = ex2f(fmaf(v.z, kLog2e, -mb)); v.w = ex2f(fmaf(v.w, kLog2e, -mb));
          store_chunk(g, q, scale4(v, sc));
        }
      }
    }
    __syncwarp();
    if (act) {
      const float* yr = reinterpret_cast<const float*>(g.srow) + g.off4;
      float* prow = pt + seq * G::PSEQ + ti * PS;
      float yv[NSL];
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const int l = label(j);
        yv[j] = yr[l >= 0 ? (l & kLabMask) : 0];
      }
      bool low = false;
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        prow[li + j * LPR] = emission(yv[j], label(j));
        low |= label(j) >= 0 && yv[j] * winv < kPMin;
      }
      // (rare) a label more than 83 nats below its row's maximum: the repair kernel redoes the sequence in the log domain
      if (low && cfg.floor_flag != nullptr && seq < gcnt) atomicOr(&cfg.floor_flag[b0 + seq], 1);
    }
  }

  // ---------------------------------------------------------------- phase 2 ahead: emissions from the slab again
  __device__ __forceinline__ void emit_step(int t, unsigned char* tsl, float* pt) const {
    if (t < Tb) {
      const RowGeom g = geom(t, tsl);
      const float* yr = reinterpret_cast<const float*>(g.srow) + g.off4;
      float* prow = pt + seq * G::PSEQ + ti * PS;
      float yv[NSL];
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const int l = label(j);
        yv[j] = yr[l >= 0 ? (l & kLabMask) : 0];
      }
#pragma unroll
      for (int j = 0; j < NSL; ++j) prow[li + j * LPR] = emission(yv[j], label(j));
    }
  }

  // ---------------------------------------------------------------- phase 2 behind: gamma scatter into the slab
  // abt: alpha/beta tiles of the item ([GB][ABSEQ]); their product is w*gamma (chain_phase2).
  // States that share a class are spread over rounds by their duplicate rank, so every round is a conflict-free
  // read-add-write on the row.
  __device__ __forceinline__ void scatter_step(int t, unsigned char* tsl, const double* abt) const {
    const bool live = t < Tb && wgt != 0.f;
    const RowGeom g = geom(t, tsl);
    float* yr = reinterpret_cast<float*>(g.srow) + g.off4;
    float gam[NSL];
    if (live) {
      const double* at = abt + seq * G::ABSEQ + ti * AS;
      const double* bt = at + TT * AS;
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const int st = li + j * LPR;
        gam[j] = label(j) >= 0 ? -(float)(at[st] * bt[st]) : 0.f;  // the row warps ADD gamma' = -w*gamma to the softmax
      }
    }
    const int nr = __reduce_max_sync(0xffffffffu, live ? max_rank : 0);
    for (int r = 0; r <= nr; ++r) {
      if (live) {
        // the addresses of one round are pairwise distinct: all loads first, then all stores
        float cur[NSL];
#pragma unroll
        for (int j = 0; j < NSL; ++j) {
          const int l = label(j);
          cur[j] = (l >= 0 && (l >> kLabBits) == r) ? yr[l & kLabMask] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < NSL; ++j) {
          const int l = label(j);
          if (l >= 0 && (l >> kLabBits) == r) yr[l & kLabMask] = cur[j] + gam[j];
        }
      }
      if (r < nr) __syncwarp();
    }
  }
};

// ============================================================================ kernel
template <int NS, int LPR, int CPL, int MINB>
__global__ void __launch_bounds__(Geo<NS, LPR>::NTHREADS, MINB) nbctc_stream_kernel(const Problem P, const StreamCfg cfg) {
  using G = Geo<NS, LPR>;
  constexpr int TT = G::TT, Lpad = G::Lpad, PS = G::PS, AS = G::AS, GB = G::GB, NSL = G::NSL, NSLOT = kNSlot;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem S;
  S.sfull = reinterpret_cast<uint64_t*>(smem_raw + cfg.o_bar);
  S.info = reinterpret_cast<int*>(smem_raw + cfg.o_info);
  S.lab = reinterpret_cast<int*>(smem_raw + cfg.o_lab);
  S.ckpt = reinterpret_cast<double*>(smem_raw + cfg.o_ckpt);
  S.cke = reinterpret_cast<int*>(smem_raw + cfg.o_cke);
  S.ptile = reinterpret_cast<float*>(smem_raw + cfg.o_ptile);
  S.s2 = reinterpret_cast<double*>(smem_raw + cfg.o_s2);
  S.pub = reinterpret_cast<ChainPub*>(smem_raw + cfg.o_pub);
  S.ab = reinterpret_cast<double*>(smem_raw + cfg.o_ab);
  S.ring = smem_raw + cfg.o_ring;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b0 = (int64_t)blockIdx.x * GB;
  const int gcnt = (int)min((int64_t)GB, P.B - b0);

  // ---- per-sequence lengths, labels, validity (include/nbctc.h parity domain)
  if (tid < 2 * kMaxGB) S.info[2 * kMaxGB + tid] = 0;
  for (int i = tid; i < NSLOT * TT; i += G::NTHREADS) mbar_init(&S.sfull[i], 1);
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
    if (tid < gcnt && cfg.floor_flag != nullptr) cfg.floor_flag[b0 + tid] = 0;  // (before the barrier: row warps may raise it)
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
  PROF_DECL

  // rows beyond the group's longest input: all-zero gradient, written directly (only ragged batches get here)
  if (want_grad && Tg < P.T) {
    const int64_t n = (int64_t)gcnt * P.C;
    for (int64_t t = Tg; t < P.T; ++t) {
      float* dst = P.grad + (t * P.B + b0) * P.C;
      for (int64_t c = tid; c < n; c += G::NTHREADS) dst[c] = 0.f;
    }
  }

#ifdef NBCTC_PROF
  int prof_it_ = 0;
#define NBCTC_ITER_END()                                                                            \
  {                                                                                                 \
    const long long t_work_ = clock64();                                                            \
    PROF_SCOPE(6, __syncthreads(); prof_[5] += *reinterpret_cast<volatile int*>(S.info) & 0)         \
    if (cfg.prof != nullptr && blockIdx.x == gridDim.x / 2 && lane == 0 && prof_it_ < 160) {         \
      cfg.prof[24 + (prof_it_ * 32 + warp) * 2] = t_work_ - prof_t0_;                                \
      cfg.prof[24 + (prof_it_ * 32 + warp) * 2 + 1] = clock64() - prof_t0_;                          \
    }                                                                                               \
    ++prof_it_;                                                                                     \
  }
#else
#define NBCTC_ITER_END() __syncthreads();
#endif

  if (warp < G::NCHAIN) {
    // ======================================================================== chain warp(s) of sequence `seq`
    constexpr int W = G::W, CNS = G::CNS, NCW = G::NCW;
    const int seq = warp / NCW;
    const bool beta_warp = NCW == 2 && (warp % NCW) == 1;  // W = 32: this warp runs beta only
    const bool isb = NCW == 2 ? beta_warp : lane >= 16;
    const int Tb = S.info[seq], Lb = S.info[kMaxGB + seq];
    const int NTb = (Tb + TT - 1) / TT;
    const float wgt = (seq < gcnt) ? P.w_scalar * (P.seq_w ? P.seq_w[b0 + seq] : 1.f) : 0.f;
    double* ck = (cfg.ckpt_global ? cfg.ws_ckpt + ((size_t)min(b0 + seq, P.B - 1) * cfg.NTmax) * Lpad
                                  : S.ckpt + ((size_t)seq * cfg.NTmax) * Lpad) + (lane & (W - 1));  // [NTmax][CNS][W]
    int* cke = cfg.ckpt_global ? cfg.ws_cke + (size_t)min(b0 + seq, P.B - 1) * cfg.NTmax * W : S.cke + (size_t)seq * cfg.NTmax * W;  // [NTmax][W]
    ChainPub* pub = S.pub + seq;
    ChainScal chain;
    double cx[CNS];
#pragma unroll
    for (int j = 0; j < CNS; ++j) cx[j] = 0.0;
    chain.carry = (lane == 0) ? 1.0 : 0.0;
    chain.zinv = 0.0;
    chain.e = 0; chain.Ez = 0;
    chain.fac = (lane & (W - 1)) == 0 ? 0.0 : 1.0;
    // ---- phase 1
    for (int it = -1; it < NTg; ++it) {
      if (!beta_warp && it >= 0 && it < NTb) {
        const float* pt = S.ptile + (size_t)((it & 1) * GB + seq) * G::PSEQ;
        PROF_SCOPE(0, chain_phase1<CNS, W, TT, PS>(cx, chain, lane, Tb, ck, cke, it, pt))
        // the gradient weight of a sequence with w = 0 is applied at the end of the kernel (see below)
        if (it == NTb - 1) chain_readout<CNS, W>(cx, chain, lane, Lb, &P.loss[b0 + seq], wgt, pub);
      }
      NBCTC_ITER_END()
    }
    // ---- phase 2: item i = tile NTg-1-i
    if (want_grad) {
      double ckv[CNS];
      int eck = 0;
      // checkpoint of tile k (k = 0 starts from the virtual state instead): the alpha direction's states and lane
      // scale; the beta direction takes the scale of the alpha lane that holds the same states
      const int ecol = isb ? W - 1 - (lane & (W - 1)) : (lane & (W - 1));
      auto fetch_ckpt = [&](int k) {
        if (k > 0 && k < NTb) {
          if (!beta_warp) {
#pragma unroll
            for (int j = 0; j < CNS; ++j) ckv[j] = ck[(k * CNS + j) * W];
          }
          eck = cke[k * W + ecol];
        } else {
#pragma unroll
          for (int j = 0; j < CNS; ++j) ckv[j] = 0.0;
          eck = 0;
        }
      };
#pragma unroll
      for (int j = 0; j < CNS; ++j) ckv[j] = 0.0;
      fetch_ckpt(NTg - 1);
      for (int i = -1; i <= NTg + 1; ++i) {
        if (i >= 0 && i < NTg) {
          const int k = NTg - 1 - i;
          if (k < NTb) {
            if (beta_warp && k == NTb - 1) {  // first tile of this sequence: take over from the alpha warp
              chain.zinv = pub->zinv;
              chain.Ez = pub->Ez;
              chain_beta_init<CNS, W>(cx, chain, lane, Lb);
            }
            double ckc[CNS];
#pragma unroll
            for (int j = 0; j < CNS; ++j) ckc[j] = ckv[j];
            const int ecc = eck;
            fetch_ckpt(k - 1);  // in flight while this tile runs
            const int buf = i & 1;
            PROF_SCOPE(1, chain_phase2<CNS, W, TT, PS, AS>(cx, chain, lane, isb, Tb, ckc, ecc, k,
                                                           S.ptile + (size_t)(buf * GB + seq) * G::PSEQ,
                                                           S.ab + (size_t)(buf * GB + seq) * G::ABSEQ))
          } else {
            fetch_ckpt(k - 1);
          }
        }
        NBCTC_ITER_END()
      }
    }
    PROF_DUMP(0)
  } else if (warp < G::NCHAIN + G::NRW) {
    // ======================================================================== row warp of time step `ti`
    const int ti = warp - G::NCHAIN;
    const int seq = lane / LPR;
    const float wgt = (seq < gcnt) ? P.w_scalar * (P.seq_w ? P.seq_w[b0 + seq] : 1.f) : 0.f;
    const Rows<NS, LPR, CPL> rows(P, cfg, S, lane, ti, gcnt, b0, wgt);
    // ---- phase 1: item a = tile a, this warp's time step t = a*TT + ti (only t < Tg is ever moved)
    uint32_t par = 0;  // bit s = parity of the next completion of this warp's barrier of slot s
    int slot_a = 0;    // ring slot of the item the ahead stage works on (ring position = items since the start)
    for (int it = -1; it < NTg; ++it) {
      const int a = it + 1;
      const int t = a * TT + ti;
      if (a < NTg) {
        if (t < Tg) {
          PROF_SCOPE(0, mbar_wait(&S.sfull[slot_a * TT + ti], (par >> slot_a) & 1u))
          par ^= 1u << slot_a;
          PROF_SCOPE(1, rows.forward_step(t, rows.slab(slot_a), S.ptile + (size_t)((a & 1) * GB) * G::PSEQ))
          if (want_grad) fence_proxy_async();  // the slab is read by the async proxy (bulk store) after the barrier
        }
        slot_a = slot_a + 1 == NSLOT ? 0 : slot_a + 1;
      }
      NBCTC_ITER_END()
    }
    if (want_grad) {
      // ---- phase 2: item i = tile NTg-1-i at ring position NTg + i; slabs come back from the gradient tensor
      int slot_b = slot_a;          // ring slot of the behind stage's item (two items behind the ahead stage)
      for (int i = -1; i <= NTg + 1; ++i) {
        const int ia = i + 1;  // ahead item: emissions
        if (ia < NTg) {
          const int ta = (NTg - 1 - ia) * TT + ti;
          if (ta < Tg) {
            PROF_SCOPE(0, mbar_wait(&S.sfull[slot_a * TT + ti], (par >> slot_a) & 1u))
            par ^= 1u << slot_a;
            PROF_SCOPE(4, rows.emit_step(ta, rows.slab(slot_a), S.ptile + (size_t)((ia & 1) * GB) * G::PSEQ))
          }
          slot_a = slot_a + 1 == NSLOT ? 0 : slot_a + 1;
        }
        const int ib = i - 1;  // behind item: gamma scatter into the slab
        if (ib >= 0 && ib < NTg) {
          const int tb = (NTg - 1 - ib) * TT + ti;
          if (tb < Tg) {
            PROF_SCOPE(5, rows.scatter_step(tb, rows.slab(slot_b), S.ab + (size_t)((ib & 1) * GB) * G::ABSEQ))
            fence_proxy_async();
          }
          slot_b = slot_b + 1 == NSLOT ? 0 : slot_b + 1;
        }
        NBCTC_ITER_END()
      }
    }
    PROF_DUMP(1)
  } else {
    // ======================================================================== mover warp: TMA for TT/NMW time steps
    // Lane l moves time step ti of every tile: loads two items ahead of the row warps, stores one iteration
    // behind them (after the barrier that ends the row warp's work on the slab).  L2 policies: logits are read
    // once (evict_first); the phase-1 gradient rows must survive in L2 until phase 2 (evict_last); phase 2 reads
    // and rewrites them for the last time (evict_first).
    constexpr int PER = TT / G::NMW;
    const int ti = (warp - G::NCHAIN - G::NRW) * PER + lane;
    const bool mine = lane < PER;
    const Rows<NS, LPR, CPL> mv(P, cfg, S, lane, mine ? ti : 0, gcnt, b0, 0.f);
    const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
    if (mine) {
      for (int a = 0; a < 3 && a < NTg; ++a)
        if (a * TT + ti < Tg) mv.issue_load(a % NSLOT, a * TT + ti, P.logits, pol_stream);
    }
    for (int it = -1; it < NTg; ++it) {
      if (mine) {
        // item `it` was finished by its row warp in the previous iteration
        if (want_grad && it >= 0) {
          if (it * TT + ti < Tg) PROF_SCOPE(0, mv.issue_store(it % NSLOT, it * TT + ti, pol_keep))
          else bulk_commit();
        }
        // slot of item it+3 was last used by item it-3, stored three iterations ago
        if (it >= 0 && it + 3 < NTg && (it + 3) * TT + ti < Tg) {
          if (want_grad) PROF_SCOPE(1, bulk_wait_read<2>())
          PROF_SCOPE(2, mv.issue_load((it + 3) % NSLOT, (it + 3) * TT + ti, P.logits, pol_stream))
        }
      }
      NBCTC_ITER_END()
    }
    if (want_grad) {
      auto t_of = [&](int i) { return (NTg - 1 - i) * TT + ti; };
      if (mine) {
        bulk_wait_all();  // every gradient row of phase 1 is in global memory before it is read back
        for (int i = 0; i < 2 && i < NTg; ++i)
          if (t_of(i) < Tg) mv.issue_load((NTg + i) % NSLOT, t_of(i), P.grad, pol_stream);
      }
      for (int i = -1; i <= NTg + 1; ++i) {
        if (mine) {
          const int is = i - 2;  // item whose slab got its gamma in the previous iteration
          if (is >= 0) {
            if (t_of(is) < Tg) PROF_SCOPE(0, mv.issue_store((NTg + is) % NSLOT, t_of(is), pol_stream))
            else bulk_commit();
          }
          // slot of item i+2 was last used by item i-4, stored two iterations ago
          if (i >= 0 && i + 2 < NTg && t_of(i + 2) < Tg) {
            PROF_SCOPE(1, bulk_wait_read<2>())
            PROF_SCOPE(2, mv.issue_load((NTg + i + 2) % NSLOT, t_of(i + 2), P.grad, pol_stream))
          }
        }
        NBCTC_ITER_END()
      }
      if (mine) bulk_wait_all();  // the ring must outlive the last bulk store's reads; rows final before the tail
    }
    PROF_DUMP(2)
  }
#undef NBCTC_ITER_END

  // Sequences with gradient weight 0 ran with weight 1 (their emissions are read back from the gradient rows):
  // their rows become zeros now.  (Their gamma scatter was skipped.)
  __syncthreads();
  if (want_grad) {
    for (int r = 0; r < gcnt; ++r) {
      const float w_r = P.w_scalar * (P.seq_w ? P.seq_w[b0 + r] : 1.f);
      const int Tr = S.info[r];
      if (w_r == 0.f && Tr > 0) {
        for (int t = 0; t < Tr; ++t) {
          float* dst = P.grad + ((int64_t)t * P.B + b0 + r) * P.C;
          for (int c = tid; c < P.C; c += G::NTHREADS) dst[c] = 0.f;
        }
      }
    }
  }
}

template <int NS, int LPR, int CPL>
int launch_inst(const Problem& p, const StreamCfg& cfg, cudaStream_t stream) {
  using G = Geo<NS, LPR>;
  const unsigned groups = (unsigned)((p.B + G::GB - 1) / G::GB);
  // registers: MINB = CTAs that should share an SM (3 for the small-tile Lpad = 256 geometry, NBCTC_CTAS otherwise)
  constexpr int kWant = NS >= 16 ? 3 : 2;
  if constexpr (G::NTHREADS * kWant <= 1024) {
    if (cfg.ctas_per_sm >= 2 || NS >= 16) {
      auto kern = nbctc_stream_kernel<NS, LPR, CPL, kWant>;
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
