// Pipeline no-blank CTC forward+backward kernel for sm_100a (single-label variant; overview in nbctc_pipe.cu).
//
// ONE persistent kernel, two warp roles, no CTA-wide barrier after start-up.  Work is handed out by two global
// ticket counters; dependencies are per sequence GROUP (GB batch-adjacent sequences whose rows at one time step
// are contiguous in the (T,B,C) tensor: a "slab") and travel through release/acquire counters in the workspace.
//
//   stage A  row warps    task = (group, block of TB time steps): slab by TMA bulk copy -> row log-partition
//                         (NoBlankCTC.py:136) -> per-state emissions p_t(s) = softmax(x_t)[label_s]
//                         (NoBlankCTC.py:96-102) and the row statistics -> aux tile (global, lives in L2)
//   stage B  chain warps  task = sequence: alpha from t = 0 upwards in lanes 0-15 and beta from t = T_b-1
//                         downwards in lanes 16-31 of ONE warp, the same instructions (NoBlankCTC.py:71-87; the
//                         reference's own beta pass is commented out at :113-125, autograd does it).  They MEET IN
//                         THE MIDDLE: the first half of the steps stores its states, the second half multiplies
//                         the fresh state of one direction with the stored state of the other:
//                         gamma_t(s) = alpha_t(s) beta_t(s) / Z -- every lattice column is computed once per
//                         direction, Z = sum_s alpha beta at the meeting point.  float64, linear domain, exact
//                         power-of-two rescaling once per 8 steps with an exponent PER LANE (neighbouring lanes
//                         exchange their scale), so states that differ by thousands of binary orders of
//                         magnitude across the lattice keep full precision.  -w*gamma overwrites p in the aux tile.
//   stage C  row warps    task = (group, block of TB time steps): the slab again (L2 hit while the group's window
//                         is resident, HBM otherwise) + the aux row by TMA, w*softmax(x) recomputed from the row
//                         statistics in place, -w*gamma scattered into it (repeated labels accumulate in
//                         duplicate-rank rounds: deterministic), slab -> gradient by TMA store.
//
// Row tasks = (group, block of TB time steps) are dealt to the CTAs round-robin; each row warp walks its CTA's task
// sequence twice (stage A ahead, stage C behind, stage C first whenever its chains are done): see row_warp_main.
// Every dependency of a task belongs to tasks at earlier positions, and waiting never blocks the consumption of
// slabs that already landed: no deadlock once every CTA of the grid runs.
//
// Template parameters: NS (chain states per lane, Lmax <= 16*NS), LPR (lanes per row; GB = 32/LPR sequences per
// group), CPL (16-byte chunks per lane and row).
#pragma once

#include <cuda_runtime.h>

#include "stream_kernel.cuh"  // PTX helpers (mbarrier, bulk copies, policies)

namespace nbctc {

constexpr int kPipeChainWarps = 4;
// Warp roles are whole warpgroups with their own register budget (setmaxnreg).  The CTA's register pool is what the
// launch allocated (threads x registers of the launch bound), so rows*32*row_regs + 4*32*chain_regs must fit in it:
//   NS <= 4:  640 threads x 96:  16 row warps x 80 + 4 chain warps x 152
//   NS >= 8:  512 threads x 128: 12 row warps x 88 + 4 chain warps x 232
template <int NS>
struct PipeTraits {
  static constexpr int kRowWarps = NS <= 4 ? 16 : 12;
  static constexpr int kRowRegs = NS <= 4 ? 80 : 88;
  static constexpr int kChainRegs = NS <= 4 ? 152 : 232;
  static constexpr int kThreads = 32 * (kRowWarps + kPipeChainWarps);
};
inline int pipe_row_warps(int ns) { return ns <= 4 ? 16 : 12; }

struct PipeCfg {
  int NS, Lpad, LPR, CPL, GB;
  int RSg;      // bytes of a slab in a ring slot: round16(GB*C*4) + 32
  int AUXF;     // floats per aux row (one group, one time step): GB*Lpad emissions / gammas + GB float2 row statistics
  int SLOTB;    // ring slot bytes: slab, aux row, label buffer (GB*Lpad labels + GB headers)
  int D;        // ring slots per row warp
  int NRW;      // row warps per CTA
  int TB;       // time steps per row task (a multiple of NRW: warp w owns steps w, w + NRW, ...)
  int TPG;      // row tasks per group = ceil(T / TB)
  int NG;       // groups = ceil(B / GB)
  int NGS;      // aux/ab/ex slots (groups in flight)
  int wmax;     // stage A runs at most `wmax` positions of a CTA's ticket sequence ahead of stage C
  int phase_mask;  // bit 0 stage A, bit 1 stage B, bit 2 stage C (all set in the fused launch)
  int want_grad;
  int keep_logits;  // L2 policy of the stage-A logits read: 0 evict_first, 1 normal, 2 evict_last
  int grid;
  uint32_t o_bar, o_meta, o_flags, o_ring, o_cring, smem_bytes;
  // workspace
  int* ctr;      // [0] (unused), [1] chain ticket
  float* zeros;  // 256 bytes of zeros
  int4* hdr;     // [B] {T_b, L_b, largest duplicate rank, gradient weight bits}; T_b = 0 outside the parity domain
  int2* grp;     // [NG] {longest T_b of the group, sequences in the group}
  int* lab;      // [B][Lpad] class | duplicate rank << 22
  int* doneA;    // [NG] finished stage-A slabs (complete = the group's longest T_b)
  int* doneB;    // [NG] finished chains
  int* doneC;    // [NG] finished stage-C slabs
  float* aux;    // [NGS][T + 2 kPadRows][AUXF]
  uint32_t* ab;  // [NGS][GB][T + 2 kPadRows][16 lanes][NSP] packed chain states + lane scales
};

int launch_pipe_ns2(const Problem& p, const PipeCfg& cfg, cudaStream_t stream);
int launch_pipe_ns4(const Problem& p, const PipeCfg& cfg, cudaStream_t stream);
int launch_pipe_ns8(const Problem& p, const PipeCfg& cfg, cudaStream_t stream);
int launch_pipe_ns16(const Problem& p, const PipeCfg& cfg, cudaStream_t stream);
int launch_pipe_prep(const Problem& p, const PipeCfg& cfg, cudaStream_t stream);

#ifdef __CUDACC__
namespace pipe {

using namespace stream;

constexpr int kRB = 8;             // chain steps between two rescales
constexpr int kPadRows = 16;       // padding rows before and after every aux / stored-state tile (chain prefetch)

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// orders generic-proxy accesses (st.global / ld.global of other threads, made visible by an acquire) with
// async-proxy accesses (TMA reads of the same global memory) of this thread
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// a dependency that never arrives is a protocol bug: trap after ~4 s instead of hanging the GPU
static __device__ __noinline__ void wait_timeout(const char* what, int a, int b) {
  printf("nbctc pipe: %s timed out (block %d warp %d: %d %d)\n", what, (int)blockIdx.x, (int)(threadIdx.x >> 5), a, b);
  __trap();
}

// ============================================================================ stage B: chain warp
template <int NS>
struct CG {
  static constexpr int Lpad = 16 * NS;
  static constexpr int U = 32 / NS;             // emission prefetch depth (steps) = unroll
  static constexpr int UO = U >= 2 ? U / 2 : 1;  // prefetch depth of the other direction's stored states
  // Largest scale step between neighbouring lanes.  Mass crosses at most ceil(8/NS) lanes between two rescales and
  // gains 2^DEC of scaled magnitude per crossing at worst: ceil(8/NS)*DEC stays below the float64 range.
  static constexpr int DEC = NS == 2 ? 208 : NS == 4 ? 420 : 850;
};

// Predicated vector loads / stores through L2 (ld/st.global.cg): the destination registers are zeroed and then
// loaded under a predicate, in straight-line code, so that a load can stay in flight for many chain steps (a
// branch or a select around the load would make the consumer wait for it at once).
__device__ __forceinline__ void ldcg2f(const float* p, bool pred, float& a, float& b) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
      "@q ld.global.cg.v2.f32 {%0, %1}, [%2];\n\t}"
      : "=f"(a), "=f"(b)
      : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ void ldcg4f(const float* p, bool pred, float& a, float& b, float& c, float& d) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
      "mov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t@q ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "=f"(a), "=f"(b), "=f"(c), "=f"(d)
      : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ void ldcg2d(const double* p, bool pred, double& a, double& b) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
      "@q ld.global.cg.v2.f64 {%0, %1}, [%2];\n\t}"
      : "=d"(a), "=d"(b)
      : "l"(p), "r"((int)pred));
}
// keeps the old value when the predicate is false
__device__ __forceinline__ void ldcg1i_keep(const int* p, bool pred, int& a) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.global.cg.s32 %0, [%1];\n\t}" : "+r"(a) : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ void stcg2f(float* p, bool pred, float a, float b) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q st.global.cg.v2.f32 [%0], {%1, %2};\n\t}" ::"l"(p), "f"(a), "f"(b), "r"((int)pred)
               : "memory");
}
__device__ __forceinline__ void stcg4f(float* p, bool pred, float a, float b, float c, float d) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\t@q st.global.cg.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(p), "f"(a), "f"(b),
               "f"(c), "f"(d), "r"((int)pred)
               : "memory");
}
__device__ __forceinline__ void stcg2d(double* p, bool pred, double a, double b) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q st.global.cg.v2.f64 [%0], {%1, %2};\n\t}" ::"l"(p), "d"(a), "d"(b), "r"((int)pred)
               : "memory");
}
__device__ __forceinline__ void stcg1i(int* p, bool pred, int a) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q st.global.cg.s32 [%0], %1;\n\t}" ::"l"(p), "r"(a), "r"((int)pred) : "memory");
}
template <int NS>
__device__ __forceinline__ void ldcg_vec(const float* p, bool pred, float (&v)[NS]) {
  if constexpr (NS == 2) {
    ldcg2f(p, pred, v[0], v[1]);
  } else {
#pragma unroll
    for (int i = 0; i < NS; i += 4) ldcg4f(p + i, pred, v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
}
template <int NS>
__device__ __forceinline__ void stcg_vec(float* p, bool pred, const float (&v)[NS]) {
  if constexpr (NS == 2) {
    stcg2f(p, pred, v[0], v[1]);
  } else {
#pragma unroll
    for (int i = 0; i < NS; i += 4) stcg4f(p + i, pred, v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
}
template <int NS>
__device__ __forceinline__ void ldcg_vec(const double* p, bool pred, double (&v)[NS]) {
#pragma unroll
  for (int i = 0; i < NS; i += 2) ldcg2d(p + i, pred, v[i], v[i + 1]);
}
template <int NS>
__device__ __forceinline__ void stcg_vec(double* p, bool pred, const double (&v)[NS]) {
#pragma unroll
  for (int i = 0; i < NS; i += 2) stcg2d(p + i, pred, v[i], v[i + 1]);
}

// x(s) <- (x(s) + x(s-1)) p(s) in the lane's position order, as fma(x(s-1), p(s), x(s) p(s)): the dependent path
// of a step is one 64-bit shuffle and one DFMA.  `fac` = 2^(e_neighbour - e_lane) brings the neighbour lane's
// last state into this lane's scale (0 for the first lane of a direction, whose shuffle returns its own value).
// kSum: sum(s) = x(s) + x(s-1) before the step (beta_t(s) resp. alpha_t(s)/p_t(s)), off the dependent path.
template <int NS, bool kSum>
__device__ __forceinline__ void chain_step(double (&x)[NS], double (&sum)[NS], const double (&p)[NS], double fac) {
  double t[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) t[j] = x[j] * p[j];
  const double up = __shfl_up_sync(0xffffffffu, x[NS - 1], 1, 16);
  const double p0f = p[0] * fac;
  if (kSum) {
#pragma unroll
    for (int j = NS - 1; j >= 1; --j) sum[j] = x[j] + x[j - 1];
    sum[0] = fma(up, fac, x[0]);
  }
#pragma unroll
  for (int j = NS - 1; j >= 1; --j) x[j] = fma(x[j - 1], p[j], t[j]);
  x[0] = fma(up, p0f, t[0]);
}

// rescale at the start of a block of 8 steps (stream_kernel.cuh: lane_rescale)
template <int NS>
__device__ __forceinline__ void block_entry(double (&x)[NS], int& e, double& fac, int hl) {
  lane_rescale<NS, 16, CG<NS>::DEC>(x, e, fac, hl);
}

struct ChainArgs {
  float* paux;      // emissions / gammas of this sequence at t = 0: + t*pstride + state (kPadRows rows of padding
                    // before t = 0 and after t = T-1: prefetches beyond the sequence stay inside the tile)
  int64_t pstride;  // floats between time steps
  uint32_t* ab;     // stored states at t = 0: + t*16*NSP; a row = 16 lanes x {NS packed states, lane scale, padding}
  int Tb, Lb;
  float wgt;
  int want_grad;
  float* loss_out;
  uint32_t ring;    // shared-memory ring of this warp (byte address in the shared window)
  const float* zeros;  // >= 64 bytes of zeros (emissions of chain lanes that hold no state)
};

// Stored chain states: a non-negative double as 32 bits = 11-bit exponent + 21-bit mantissa (round to nearest).
// gamma only needs ~1e-7 relative accuracy of the stored factor; the recursions themselves stay in float64.
__device__ __forceinline__ uint32_t pack_state(double x) {
  return (uint32_t)((unsigned long long)(__double_as_longlong(x) + (1ll << 30)) >> 31);
}
__device__ __forceinline__ double unpack_state(uint32_t u) { return __longlong_as_double((long long)((unsigned long long)u << 31)); }

template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src) {
  if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NS>
struct CRing {
  static constexpr int NSP = NS + (NS == 2 ? 2 : 4);      // uint32 per lane and step of stored state: states, scale, padding
  static constexpr int PB = NS * 4, OB = NSP * 4;         // bytes per lane and step: emissions, stored state
  static constexpr int STEP = 32 * (PB + OB);             // ring bytes per step
  static constexpr int U = NS == 2 ? 12 : NS == 4 ? 8 : NS == 8 ? 5 : 3;  // steps in flight
  static constexpr int BYTES = U * STEP;
};
constexpr int pipe_chain_ring_bytes(int ns) {
  return ns == 2 ? CRing<2>::BYTES : ns == 4 ? CRing<4>::BYTES : ns == 8 ? CRing<8>::BYTES : CRing<16>::BYTES;
}

// One sequence on one warp.  Iteration k of the FIRST half: alpha (lanes 0-15) at t = k, beta (lanes 16-31) at
// t = T_b-1-k+odd; both store their state and lane scale.  Iteration k of the SECOND half: beta at t = Ha-1-k, alpha
// at t = Ha+k; gamma_t = (sum of the fresh direction before its step) * (state the other direction stored at t) / Z.
// T_b odd: the beta direction starts with a virtual step at t = T_b whose emission is 1 at the start state, which
// leaves the start vector unchanged; the alpha direction's last iteration then has nothing to do.
// Every lane fetches its own emissions and stored states with cp.async into a shared-memory ring, U steps ahead
// (cp.async.wait_group keeps exactly U-1 steps in flight, which register prefetching cannot: the hardware has six
// scoreboards); the copies are unconditional -- padding rows keep the addresses valid -- and stores are predicated,
// so a body of 8 steps is branch-free.
template <int NS>
struct ChainRun {
  using G = CG<NS>;
  using RG = CRing<NS>;
  static constexpr int Lpad = G::Lpad, U = RG::U, NSP = RG::NSP;
  static constexpr int UB = kRB;              // steps per body
  static constexpr int ABS = 16 * NSP;        // uint32 per row of stored states
  const ChainArgs& a;
  const int lane, hl;
  const bool isb;
  int Ha, odd;
  bool live;
  unsigned smask;
  double x[NS], sum[NS];
  int e;
  double fac;
  int64_t pstep, gstep, abstep;
  const float* pld;
  const uint32_t* old_;
  float* gst;
  uint32_t* abw;
  uint32_t slot_addr, ring_lo, ring_hi;  // ring position of the step that is consumed / refilled next
  int Ez, kx;
  double zinv, sA, sB;

  __device__ __forceinline__ ChainRun(const ChainArgs& a_, int lane_) : a(a_), lane(lane_), hl(lane_ & 15), isb(lane_ >= 16) {}

  __device__ __forceinline__ void advance_slot() {
    slot_addr += RG::STEP;
    if (slot_addr == ring_hi) slot_addr = ring_lo;
  }
  // this lane's emissions of the next step to fetch -> ring slot at `dst`
  __device__ __forceinline__ void fetch_p(uint32_t dst) {
    if constexpr (NS == 2) {
      cp_async<8>(dst + lane * RG::PB, pld);
    } else {
#pragma unroll
      for (int i = 0; i < NS; i += 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + lane * RG::PB + i * 4), "l"(pld + i) : "memory");
    }
    pld += pstep;
  }
  __device__ __forceinline__ void fetch_o(uint32_t dst) {
#pragma unroll
    for (int i = 0; i < NSP; i += 4) cp_async<16>(dst + 32 * RG::PB + lane * RG::OB + i * 4, old_ + i);
    old_ += abstep;
  }
  // emissions of the current slot in the lane's position order; lanes without a state hold zeros
  __device__ __forceinline__ void read_p(double (&p)[NS]) const {
    float raw[NS];
    const uint32_t src = slot_addr + lane * RG::PB;
    if constexpr (NS == 2) {
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(raw[0]), "=f"(raw[1]) : "r"(src));
    } else {
#pragma unroll
      for (int i = 0; i < NS; i += 4)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(raw[i]), "=f"(raw[i + 1]), "=f"(raw[i + 2]), "=f"(raw[i + 3]) : "r"(src + i * 4));
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) p[j] = (double)(isb ? raw[NS - 1 - j] : raw[j]);  // (lanes without a state fetch zeros)
  }
  // stored states of the other direction (its lane 15-hl holds this lane's states in reversed order) and its scale
  __device__ __forceinline__ void read_o(double (&o)[NS], int& eo) const {
    uint32_t v[NSP];
    const uint32_t src = slot_addr + 32 * RG::PB + lane * RG::OB;
#pragma unroll
    for (int i = 0; i < NSP; i += 4)
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i]), "=r"(v[i + 1]), "=r"(v[i + 2]), "=r"(v[i + 3]) : "r"(src + i * 4));
#pragma unroll
    for (int q = 0; q < NS; ++q) o[q] = unpack_state(v[NS - 1 - q]);
    eo = (int)v[NS];
  }
  __device__ __forceinline__ void store_state(bool ok) {
    uint32_t v[NSP];
#pragma unroll
    for (int q = 0; q < NS; ++q) v[q] = pack_state(x[q]);
    v[NS] = (uint32_t)e;
#pragma unroll
    for (int q = NS + 1; q < NSP; ++q) v[q] = 0u;
    if (ok) {
#pragma unroll
      for (int i = 0; i < NSP; i += 4) __stcg(reinterpret_cast<uint4*>(abw + i), make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]));
    }
    abw += abstep;
  }

  // ------------------------------------------------------------------ first half
  template <bool kFirst, bool kTail>
  __device__ __forceinline__ void body1(int base) {
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const int k = base + j;
      if (j == 0) block_entry<NS>(x, e, fac, hl);
      double p[NS];
      cp_async_wait<U - 1>();
      read_p(p);
      fetch_p(slot_addr);
      cp_async_commit();
      advance_slot();
      bool st_ok = live;
      if (kFirst && j == 0) {
        // first step of alpha; first step of beta (T_b even) or its virtual step (T_b odd)
        const bool virt = isb && odd;
#pragma unroll
        for (int q = 0; q < NS; ++q) p[q] = ((smask >> q) & 1u) ? (virt ? 1.0 : p[q]) : 0.0;
        st_ok = live && !virt;
      }
      if (kFirst && j == 1) {
        const bool first_beta = isb && odd;
#pragma unroll
        for (int q = 0; q < NS; ++q) p[q] = (!first_beta || ((smask >> q) & 1u)) ? p[q] : 0.0;
      }
      if (kTail) {
        const bool act = k < Ha;
        double xs[NS];
#pragma unroll
        for (int q = 0; q < NS; ++q) xs[q] = x[q];
        chain_step<NS, false>(x, sum, p, fac);
#pragma unroll
        for (int q = 0; q < NS; ++q) x[q] = act ? x[q] : xs[q];
        st_ok = st_ok && act;
      } else {
        chain_step<NS, false>(x, sum, p, fac);
      }
      store_state(st_ok);
    }
  }

  // ------------------------------------------------------------------ second half
  template <bool kTail>
  __device__ __forceinline__ void body2(int base) {
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const int k = base + j;
      if (j == 0 && base > 0) block_entry<NS>(x, e, fac, hl);
      double p[NS], o[NS];
      int eo;
      cp_async_wait<U - 1>();
      read_p(p);
      read_o(o, eo);
      fetch_p(slot_addr);
      fetch_o(slot_addr);
      cp_async_commit();
      advance_slot();
      // gamma = sum * stored * 2^(e + eo - Ez) / zhat; the power of two is split over both factors (range).  e changes
      // at this direction's block entries (j = 0), eo when the other direction's block changes (j = kx)
      if (j == 0 || j == kx) {
        const int dd = e + eo - Ez;
        sA = pow2z(dd >> 1);
        sB = pow2z(dd - (dd >> 1)) * zinv;
      }
      chain_step<NS, true>(x, sum, p, fac);
      float g[NS], gm[NS];
#pragma unroll
      for (int q = 0; q < NS; ++q) g[q] = (float)((sum[q] * sA) * (o[q] * sB));
#pragma unroll
      for (int q = 0; q < NS; ++q) gm[q] = isb ? g[NS - 1 - q] : g[q];
      // (alpha's last iteration is beyond T_b when T_b is odd)
      stcg_vec<NS>(gst, live && (!kTail || k < Ha) && (isb || k < Ha - odd), gm);
      gst += gstep;
    }
  }

  __device__ __forceinline__ void run() {
    const int Tb = a.Tb, Lb = a.Lb;
    Ha = (Tb + 1) >> 1;
    odd = Tb & 1;
    // position q = hl*NS + j of a direction is state q for alpha and state Lpad-1-q for beta: both shift the same way
    const int soff = isb ? Lpad - NS - hl * NS : hl * NS;  // the lane's states in memory order start here
    live = soff < Lb;  // (stage A wrote p = 0 for the padded states of a live lane)
    smask = 0;         // bit j: state is the direction's start state
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int sj = isb ? Lpad - 1 - (hl * NS + j) : hl * NS + j;
      if (sj == (isb ? Lb - 1 : 0)) smask |= 1u << j;
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) x[j] = ((smask >> j) & 1u) ? 1.0 : 0.0;
    e = 0;
    fac = hl == 0 ? 0.0 : 1.0;
    // alpha walks up in time, beta walks down (both halves); lanes without a state read a block of zeros
    pstep = !live ? 0 : isb ? -a.pstride : a.pstride;
    gstep = isb ? -a.pstride : a.pstride;
    abstep = isb ? -(int64_t)ABS : (int64_t)ABS;
    ring_lo = a.ring;
    ring_hi = a.ring + RG::BYTES;
    slot_addr = ring_lo;

    // ---- first half (beta's iteration 0 is virtual when T_b is odd: its row T_b is padding or unused)
    const int t1_0 = isb ? Tb - 1 + odd : 0;
    pld = live ? a.paux + soff + (int64_t)t1_0 * a.pstride : a.zeros;
    for (int j = 0; j < U; ++j) {
      fetch_p(ring_lo + j * RG::STEP);
      cp_async_commit();
    }
    abw = a.ab + (int64_t)t1_0 * ABS + hl * NSP;
    {
      int base = 0;
      if (UB <= Ha) {
        body1<true, false>(0);
        for (base = UB; base + UB <= Ha; base += UB) body1<false, false>(base);
        if (base < Ha) body1<false, true>(base);
      } else {
        body1<true, true>(0);
      }
    }
    cp_async_wait<0>();
    __syncwarp();  // the other half-warp's stores are visible to this lane's copies below

    // ---- second half
    const int t2_0 = isb ? Ha - 1 : Ha;
    old_ = a.ab + (int64_t)t2_0 * ABS + (15 - hl) * NSP;
    gst = a.paux + soff + (int64_t)t2_0 * a.pstride;  // gamma' overwrites the emissions of the same time step
    pld = live ? gst : a.zeros;
    slot_addr = ring_lo;
    for (int j = 0; j < U; ++j) {
      fetch_p(ring_lo + j * RG::STEP);
      fetch_o(ring_lo + j * RG::STEP);
      cp_async_commit();
    }

    block_entry<NS>(x, e, fac, hl);
    // ---- Z = sum_s alpha_t(s) beta_t(s) at t = Ha-1 (beta lanes): beta_t(s) = sum of the beta direction's next step
    {
      const double up = __shfl_up_sync(0xffffffffu, x[NS - 1], 1, 16);
#pragma unroll
      for (int j = NS - 1; j >= 1; --j) sum[j] = x[j] + x[j - 1];
      sum[0] = fma(up, fac, x[0]);
      double o[NS];
      int eo0;
      cp_async_wait<U - 1>();
      read_o(o, eo0);
      const int ms = live ? top_exponent<NS>(sum) : kSent, mo = top_exponent<NS>(o);
      double part = 0.0;
      int ep = kSent;
      if (ms != kSent && mo != kSent) {
        const double s1 = pow2z(-ms), s2 = pow2z(-mo);
#pragma unroll
        for (int j = 0; j < NS; ++j) part = fma(sum[j] * s1, o[j] * s2, part);
        if (part > 0.0) ep = e + eo0 + ms + mo;
      }
      int emax = ep;
#pragma unroll
      for (int o2 = 8; o2 > 0; o2 >>= 1) emax = max(emax, __shfl_xor_sync(0xffffffffu, emax, o2, 16));
      part = ep == kSent ? 0.0 : part * pow2z(ep - emax);
#pragma unroll
      for (int o2 = 8; o2 > 0; o2 >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o2, 16);
      const double zsum = __shfl_sync(0xffffffffu, part, 16);
      emax = __shfl_sync(0xffffffffu, emax, 16);
      const int ezf = __double2hiint(zsum) >> 20;
      const bool writer = hl == 0 && !isb;
      if (emax == kSent || !(zsum > 0.0) || ezf <= 0 || ezf >= 0x7ff) {
        if (writer) *a.loss_out = INFINITY;
        zinv = 0.0;
        Ez = 0;
      } else {
        const int ezz = ezf - 1023;
        const double zhat = zsum * pow2z(-ezz);
        Ez = emax + ezz;
        if (writer) *a.loss_out = (float)(-(log(zhat) + (double)Ez * 0.6931471805599453));
        zinv = -(double)a.wgt / zhat;  // negative: stage C ADDS gamma' = -w*gamma to w*softmax
      }
    }
    if (a.want_grad) {
      kx = Ha & 7;  // the other direction's block of 8 steps changes when Ha-1-k crosses a multiple of 8
      int base = 0;
      for (; base + UB <= Ha; base += UB) body2<false>(base);
      if (base < Ha) body2<true>(base);
    }
    cp_async_wait<0>();  // nothing may land in the ring after the next sequence has started
  }
};

template <int NS>
__device__ __forceinline__ void chain_sequence(const ChainArgs& a, const int lane) {
  if (a.Tb == 1) {  // one frame, one state: gamma = 1
    if (lane == 0) {
      const float p0 = __ldcg(a.paux);
      *a.loss_out = -logf(p0);
      if (a.want_grad) __stcg(a.paux, -a.wgt);
    }
    return;
  }
  ChainRun<NS> r(a, lane);
  r.run();
}

// ============================================================================ stages A and C: row warp
// Geometry of one (t,b) row seen as 16-byte chunks of its slab (stream_kernel.cuh: RowGeom, mask_head, mask_tail,
// store_part, scale4).
template <int NS, int LPR, int CPL>
struct PRows {
  static constexpr int Lpad = 16 * NS, GB = 32 / LPR, NSL = Lpad / LPR;
  static_assert(NSL >= 1 && NSL <= 16, "states per lane of the gather/scatter");

  const Problem& P;
  const PipeCfg& cfg;
  const int lane, li, seq;
  const int C;
  // task state
  int64_t b0;
  int gcnt, Tb, Lb, max_rank, ph_fixed;
  float wgt;
  uint32_t gbytes;
  int labr[NSL];

  __device__ __forceinline__ PRows(const Problem& P_, const PipeCfg& cfg_, int lane_)
      : P(P_), cfg(cfg_), lane(lane_), li(lane_ & (LPR - 1)), seq(lane_ / LPR), C((int)P_.C), b0(0), gcnt(0), Tb(0), Lb(0),
        max_rank(0), ph_fixed(-1), wgt(0.f), gbytes(0) {}

  // group geometry only (producer side: TMA addresses)
  __device__ __forceinline__ void set_group(int g) {
    b0 = (int64_t)g * GB;
    gcnt = (int)min((int64_t)GB, P.B - b0);
    gbytes = (uint32_t)gcnt * (uint32_t)C * 4u;
    ph_fixed = (((unsigned)P.B * (unsigned)C) & 3u) == 0 ? (int)((((unsigned)b0 & 3u) * ((unsigned)C & 3u)) & 3u) : -1;
  }
  // + per-sequence state of the consumer (lengths, weight, labels of this lane's states) from the slot's label buffer:
  // [GB][Lpad] labels, then [GB] headers
  __device__ __forceinline__ void begin_task(int g, const unsigned char* lbuf) {
    set_group(g);
    Tb = 0; Lb = 0; max_rank = 0; wgt = 0.f;
    if (seq < gcnt) {
      const int4 h = *reinterpret_cast<const int4*>(lbuf + GB * Lpad * 4 + seq * 16);
      Tb = h.x; Lb = h.y; max_rank = h.z; wgt = __int_as_float(h.w);
    }
    const int* lab_seq = reinterpret_cast<const int*>(lbuf) + seq * Lpad;
#pragma unroll
    for (int j = 0; j < NSL; ++j) {
      const int st = li + j * LPR;
      labr[j] = st < Lb ? lab_seq[st] : -1;
    }
  }

  __device__ __forceinline__ uint64_t elem_off(int t) const { return (((uint64_t)t * P.B + b0) * P.C) * 4u; }
  __device__ __forceinline__ int slab_phase(int t) const {
    if (ph_fixed >= 0) return ph_fixed;
    return (int)(((((unsigned)t & 3u) * ((unsigned)P.B & 3u) + ((unsigned)b0 & 3u)) * ((unsigned)C & 3u)) & 3u);
  }
  __device__ __forceinline__ RowGeom geom(int t, unsigned char* tsl) const {
    const int fidx = slab_phase(t) + seq * C;
    RowGeom g;
    g.off4 = fidx & 3;
    g.nch = (g.off4 + C + 3) >> 2;
    g.rem = g.off4 + C - 4 * (g.nch - 1);
    g.srow = reinterpret_cast<float4*>(tsl) + (fidx >> 2);
    return g;
  }

  // ---------------------------------------------------------------- TMA (lane 0 only)
  // rows of time step t -> slab: the 16-byte aligned superset; returns nothing, arrives on `bar` with the byte count
  // (+ extra_tx bytes of a second copy the caller issues on the same barrier)
  __device__ __forceinline__ void issue_load(unsigned char* dst, uint64_t* bar, int t, uint64_t pol, const float* aux_src,
                                             uint32_t aux_bytes) const {
    const uint64_t a = reinterpret_cast<uint64_t>(P.logits) + elem_off(t);
    const uint64_t lim = reinterpret_cast<uint64_t>(P.logits) + (uint64_t)P.T * P.B * P.C * 4u;
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
    uint32_t tx = 0;
    if (a1 > a0) {
      bulk_g2s_hint(smem_u32(dst), a0, (uint32_t)(a1 - a0), smem_u32(bar), pol);
      tx += (uint32_t)(a1 - a0);
    }
    if (aux_bytes) {
      bulk_g2s_hint(smem_u32(dst + cfg.RSg), reinterpret_cast<uint64_t>(aux_src), aux_bytes, smem_u32(bar), pol);
      tx += aux_bytes;
    }
    if (tx) mbar_arrive_expect_tx(bar, tx);
    else mbar_arrive(bar);
  }
  // finished slab of time step t -> gradient rows: aligned interior as one bulk store, <= 3 floats per side by hand
  __device__ __forceinline__ void issue_store(const unsigned char* slab, int t, uint64_t pol) const {
    const uint64_t g = reinterpret_cast<uint64_t>(P.grad) + elem_off(t);
    const uint64_t gend = g + gbytes;
    uint64_t g0 = (g + 15) & ~uint64_t(15), g1 = gend & ~uint64_t(15);
    const unsigned char* src = slab + (g & 15);
    if (g1 > g0) {
      bulk_s2g_hint(g0, smem_u32(src + (g0 - g)), (uint32_t)(g1 - g0), pol);
    } else {
      g0 = gend; g1 = gend;
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
  // chunk slot c of a row is chunk q = li + c*LPR; slots c <= CPL-3 always hold a full inner chunk
  static __device__ __forceinline__ constexpr bool slot_is_inner(int c) { return c + 3 <= CPL; }
  __device__ __forceinline__ float4 load_chunk(const RowGeom& g, int q) const {
    float4 v = make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
    if (q < g.nch) {
      v = g.srow[q];
      if (q == 0) mask_head(v, g.off4);
      if (q == g.nch - 1) mask_tail(v, g.rem);
    }
    return v;
  }
  __device__ __forceinline__ float4 load_slot(const RowGeom& g, int c) const {
    const int q = li + c * LPR;
    if (slot_is_inner(c)) {
      float4 v = g.srow[q];
      if (c == 0 && li == 0) mask_head(v, g.off4);
      return v;
    }
    return load_chunk(g, q);
  }
  __device__ __forceinline__ void store_chunk(const RowGeom& g, int q, const float4& y) const {
    if (q < g.nch) {
      const int lo = q == 0 ? g.off4 : 0, hi = q == g.nch - 1 ? g.rem : 4;
      if (lo == 0 && hi == 4) g.srow[q] = y;
      else store_part(g.srow + q, y, lo, hi);
    }
  }
  __device__ __forceinline__ void store_slot(const RowGeom& g, int c, const float4& y) const {
    const int q = li + c * LPR;
    if (slot_is_inner(c) && c > 0) g.srow[q] = y;
    else store_chunk(g, q, y);
  }

  // ---------------------------------------------------------------- stage A: one slab
  // row maximum and sum of exponentials (NoBlankCTC.py:136) -> row statistics (m*log2e, 1/sum) and the emissions
  // p_t(s) = softmax(x_t)[label_s] (NoBlankCTC.py:96-102) -> aux row in global memory
  __device__ __forceinline__ void stage_a(int t, unsigned char* tsl, float* auxrow) const {
    const bool act = t < Tb;
    const RowGeom g = geom(t, tsl);
    float4 v[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) v[c] = load_slot(g, c);
    float m_l = kNegInf;
#pragma unroll
    for (int c = 0; c < CPL; ++c) m_l = fmaxf(m_l, fmaxf(fmaxf(v[c].x, v[c].y), fmaxf(v[c].z, v[c].w)));
    const float ml2 = (m_l > kNegInf) ? m_l * kLog2e : 0.f;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      s0 += ex2f(fmaf(v[c].x, kLog2e, -ml2));
      s1 += ex2f(fmaf(v[c].y, kLog2e, -ml2));
      s2 += ex2f(fmaf(v[c].z, kLog2e, -ml2));
      s3 += ex2f(fmaf(v[c].w, kLog2e, -ml2));
    }
    const float m = group_max(m_l);
    const float cf = (m_l > kNegInf) ? ex2f((m_l - m) * kLog2e) : 0.f;  // this lane's exponentials -> row maximum
    const float s = group_sum(((s0 + s1) + (s2 + s3)) * cf);
    if (act) {
      const float mb = m * kLog2e;
      const float rs = __fdividef(1.f, s);  // 1 <= s <= C
      float* arow = auxrow + seq * Lpad;
      if (li == 0) __stcg(reinterpret_cast<float2*>(auxrow + GB * Lpad + 2 * seq), make_float2(mb, rs));
      const float* yr = reinterpret_cast<const float*>(g.srow) + g.off4;
      float xv[NSL];
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const int l = labr[j];
        xv[j] = yr[l >= 0 ? (l & kLabMask) : 0];
      }
      // states L_b .. roundup(L_b, NS)-1 share a chain lane with real states: their emission is 0
      const int Lz = (Lb + NS - 1) / NS * NS;
      float* ar = arow + li;
#pragma unroll
      for (int j = 0; j < NSL; ++j) {
        const float pv = labr[j] >= 0 ? fmaxf(ex2f(fmaf(xv[j], kLog2e, -mb)) * rs, kPMin) : 0.f;
        if (li + j * LPR < Lz) __stcg(ar + j * LPR, pv);
      }
    }
  }

  // ---------------------------------------------------------------- stage C: one slab
  // slab -> w*softmax(x) in place from the stored row statistics (zeros beyond input_length),
  // -w*gamma (aux row, shared memory) added at the label classes in duplicate-rank rounds
  __device__ __forceinline__ void stage_c(int t, unsigned char* tsl, const float* auxs) const {
    const bool act = t < Tb && wgt != 0.f;
    const RowGeom g = geom(t, tsl);
    float mb = INFINITY, sc = 0.f;  // ex2(-inf) = 0: inactive rows become zeros
    if (act) {
      const float2 st = *reinterpret_cast<const float2*>(auxs + GB * Lpad + 2 * seq);
      mb = st.x;
      sc = wgt * st.y;
    }
    float4 v[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) v[c] = load_slot(g, c);
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      v[c].x = ex2f(fmaf(v[c].x, kLog2e, -mb)) * sc;
      v[c].y = ex2f(fmaf(v[c].y, kLog2e, -mb)) * sc;
      v[c].z = ex2f(fmaf(v[c].z, kLog2e, -mb)) * sc;
      v[c].w = ex2f(fmaf(v[c].w, kLog2e, -mb)) * sc;
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c) store_slot(g, c, v[c]);
    __syncwarp();
    float* yr = reinterpret_cast<float*>(g.srow) + g.off4;
    float gam[NSL];
    if (act) {
      const float* gr = auxs + seq * Lpad;
#pragma unroll
      for (int j = 0; j < NSL; ++j) gam[j] = labr[j] >= 0 ? gr[li + j * LPR] : 0.f;
    }
    const int nr = __reduce_max_sync(0xffffffffu, act ? max_rank : 0);
    for (int r = 0; r <= nr; ++r) {
      if (act) {
        // the addresses of one round are pairwise distinct: all loads first, then all stores
        float cur[NSL];
#pragma unroll
        for (int j = 0; j < NSL; ++j) {
          const int l = labr[j];
          cur[j] = (l >= 0 && (l >> kLabBits) == r) ? yr[l & kLabMask] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < NSL; ++j) {
          const int l = labr[j];
          if (l >= 0 && (l >> kLabBits) == r) yr[l & kLabMask] = cur[j] + gam[j];
        }
      }
      if (r < nr) __syncwarp();
    }
  }
};

// meta word of a ring slot
constexpr int kMetaA = 1, kMetaC = 2, kMetaFirst = 4, kMetaLast = 8;

// Row tickets are handed out statically: CTA c owns tickets c, c + grid, c + 2 grid, ...; a ticket is a task (group,
// block of TB = KS*NRW time steps) and row warp w owns its time steps w, w + NRW, ...  Every row warp walks its CTA's
// sequence twice, without talking to the other warps: position ka = next stage-A task, kc <= ka = next stage-C
// task.  Stage C goes first whenever the chains of its group are done; otherwise stage A runs ahead, at most `wmax`
// positions: the window between the stages adapts to the chain latency, and it bounds the L2 footprint (the tickets
// in flight span grid*TB slabs per stage plus the window).  Every dependency of a task belongs to tasks at earlier
// positions: no deadlock once every CTA of the grid has started.
template <int NS, int LPR, int CPL>
__device__ __forceinline__ void row_warp_main(const Problem& P, const PipeCfg& cfg, unsigned char* smem_raw, int rw, int lane) {
  using R = PRows<NS, LPR, CPL>;
  constexpr int GB = R::GB, Lpad = R::Lpad;
  const int D = cfg.D, NRW = cfg.NRW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + cfg.o_bar) + rw * D;
  int4* meta = reinterpret_cast<int4*>(smem_raw + cfg.o_meta) + rw * D;
  volatile int* flags = reinterpret_cast<volatile int*>(smem_raw + cfg.o_flags);  // dependencies known to be met
  unsigned char* ring = smem_raw + cfg.o_ring + (size_t)rw * D * cfg.SLOTB;
  R prod(P, cfg, lane), cons(P, cfg, lane);
  const bool do_a = cfg.phase_mask & 1, do_c = (cfg.phase_mask & 4) && cfg.want_grad;
  const int NA = cfg.NG * cfg.TPG;
  const int T = (int)P.T;
  const int G = (int)gridDim.x, c0 = (int)blockIdx.x;
  // L2 policy of the stage-A read of the logits: evict_last keeps them for stage C when the window of groups between
  // the stages fits into L2 (cfg.keep_logits), otherwise they are streamed like everything else
  uint64_t pol_keep;
  if (cfg.keep_logits == 2) pol_keep = policy_evict_last();
  else if (cfg.keep_logits == 1) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_keep));
  else pol_keep = policy_evict_first();
  const uint64_t pol_stream = policy_evict_first();
  const uint32_t aux_row_bytes = (uint32_t)cfg.AUXF * 4u;

  // producer cursor (warp-uniform)
  int cur_type = 0, cur_gi = 0, cur_t = 0, cur_tend = 0, cur_first = 0;
  // positions ka (stage A) and kc (stage C) of the CTA's task sequence c0, c0 + G, ... as (group, time block)
  const int Gq = G / cfg.TPG, Gr = G - Gq * cfg.TPG;
  int ka = 0, kc = 0;
  int a_gi = c0 / cfg.TPG, a_tb = c0 - a_gi * cfg.TPG;
  int c_gi = a_gi, c_tb = a_tb;
  auto advance = [&](int& gi, int& tb) {
    tb += Gr;
    gi += Gq;
    if (tb >= cfg.TPG) {
      tb -= cfg.TPG;
      ++gi;
    }
  };
  bool finished = false;
  int head = 0, tail = 0, inflight = 0, attempts = 0, task_n = 0;
  uint32_t par = 0;
  const size_t aux_slot_rows = (size_t)T + 2 * kPadRows;

  auto aux_row = [&](int gi, int t) { return cfg.aux + ((size_t)(gi & (cfg.NGS - 1)) * aux_slot_rows + kPadRows + t) * cfg.AUXF; };
  // Dependency `which` (0: the chains of a stage-C task's group, 1: the aux slot of a stage-A task) of position k.
  // All row warps of the CTA wait for the same things in the same order: what one of them has seen is cached in shared
  // memory, and only warp k % NRW polls the global counter at full rate.
  auto dep_ready = [&](int which, int k, const int* ctr, int want) -> bool {
    if (k < flags[which]) {
      __threadfence_block();
      return true;
    }
    const int poller = (k & 15) < NRW ? (k & 15) : 0;
    if (poller != rw && ((++attempts) & 7) != 0) return false;
    int ok = 0;
    if (lane == 0) ok = ld_acquire(ctr) >= want;
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (!ok) return false;
    __threadfence_block();
    if (lane == 0) atomicMax(const_cast<int*>(&flags[which]), k + 1);
    return true;
  };
  // labels + sequence headers of group gi -> label buffer of a slot (16-byte async copies, one commit group)
  auto prefetch_labels = [&](int gi, unsigned char* slot) {
    const int64_t b0 = (int64_t)gi * GB;
    const int gcnt = (int)min((int64_t)GB, P.B - b0);
    const uint32_t lb = smem_u32(slot) + cfg.RSg + aux_row_bytes;
    const int nlab = gcnt * (Lpad / 4);
    const char* src = reinterpret_cast<const char*>(cfg.lab + b0 * Lpad);
    if constexpr (GB * (Lpad / 4) <= 32) {
      if (lane < nlab) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lb + lane * 16), "l"(src + lane * 16) : "memory");
    } else {
      for (int c = lane; c < nlab; c += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lb + c * 16), "l"(src + (size_t)c * 16) : "memory");
    }
    if (lane < gcnt)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lb + GB * Lpad * 4 + lane * 16),
                   "l"(reinterpret_cast<const char*>(cfg.hdr + b0) + lane * 16)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // returns 1 if a load was issued, 0 if the producer has to wait (or is done)
  auto try_produce = [&]() -> int {
    for (;;) {
      if (cur_type == 0) {
        const bool haveA = do_a && a_gi < cfg.NG, haveC = do_c && c_gi < cfg.NG;
        if (!haveA && !haveC) {
          finished = true;
          return 0;
        }
        bool started = false;
        if (haveC && (kc < ka || !haveA)) {
          const int gi = c_gi;
          const int2 gg = __ldg(&cfg.grp[gi]);
          if (dep_ready(0, kc, cfg.doneB + gi, gg.y)) {  // the group's chains
            // gamma' and the row statistics were written through the generic proxy, the bulk copies read them
            // through the async proxy
            asm volatile("fence.proxy.async.global;" ::: "memory");
            const int Tg = gg.x;
            const int t0 = c_tb * cfg.TB + rw, t1 = min(T, (c_tb + 1) * cfg.TB), tl = min(t1, Tg);
            prod.set_group(gi);
            // rows beyond the group's longest input: zeros, written directly (ragged batches only)
            if (t1 > Tg) {
              int t = t0;
              while (t < Tg) t += NRW;
              for (; t < t1; t += NRW) {
                float* dst = P.grad + ((int64_t)t * P.B + prod.b0) * P.C;
                const int n = prod.gcnt * (int)P.C;
                for (int c = lane; c < n; c += 32) dst[c] = 0.f;
              }
            }
            if (t0 < tl) {
              cur_type = kMetaC; cur_gi = gi; cur_t = t0; cur_tend = tl; cur_first = 1;
            }
            ++kc;
            advance(c_gi, c_tb);
            started = true;
          }
        }
        if (!started && haveA && (!do_c || ka - kc < cfg.wmax)) {
          const int gi = a_gi;
          bool ok = true;
          if (gi >= cfg.NGS) {
            // the aux slot of this group was used by group gi - NGS: its last reader must be done
            const int gp = gi - cfg.NGS;
            const int2 gg = __ldg(&cfg.grp[gp]);
            ok = do_c ? dep_ready(1, ka, cfg.doneC + gp, gg.x) : dep_ready(1, ka, cfg.doneB + gp, gg.y);
          }
          if (ok) {
            const int Tg = __ldg(&cfg.grp[gi].x);
            const int t0 = a_tb * cfg.TB + rw, t1 = min(min(T, (a_tb + 1) * cfg.TB), Tg);
            if (t0 < t1) {
              cur_type = kMetaA; cur_gi = gi; cur_t = t0; cur_tend = t1; cur_first = 1;
              prod.set_group(gi);
            }
            ++ka;
            advance(a_gi, a_tb);
            started = true;
          }
        }
        if (!started) return 0;
        if (cur_type == 0) continue;  // nothing to do for this warp in that task
      }
      // issue the load of (cur_type, cur_gi, cur_t) into slot `head`
      const int t = cur_t;
      cur_t += NRW;
      const bool last = cur_t >= cur_tend;
      unsigned char* slot = ring + (size_t)head * cfg.SLOTB;
      if (cur_first) prefetch_labels(cur_gi, slot);
      if (lane == 0) {
        meta[head] = make_int4(cur_type | (cur_first ? kMetaFirst : 0) | (last ? kMetaLast : 0), t, cur_gi, 0);
        if (cur_type == kMetaA) prod.issue_load(slot, &bars[head], t, pol_keep, nullptr, 0);
        else prod.issue_load(slot, &bars[head], t, pol_stream, aux_row(cur_gi, t), aux_row_bytes);
      }
      __syncwarp();
      cur_first = 0;
      if (last) cur_type = 0;
      head = head + 1 == D ? 0 : head + 1;
      ++inflight;
      return 1;
    }
  };

  int* sig_ctr = nullptr;
  int sig_n = 0;
  auto flush_signal = [&]() {
    if (sig_ctr != nullptr) {
      __syncwarp();  // every lane's emission stores are ordered before lane 0's release
      if (lane == 0) red_release_add(sig_ctr, sig_n);
      sig_ctr = nullptr;
    }
  };
  unsigned long long t_idle = 0;
  for (;;) {
    while (inflight < D - 1) {
      if (!try_produce()) break;
    }
    if (inflight == 0) {
      flush_signal();
      if (finished) break;
      const unsigned long long now = gtime();
      if (t_idle == 0) t_idle = now;
      else if (now - t_idle > 4000000000ull) wait_timeout("row warp dependency", ka, kc);
      __nanosleep(100);
      continue;
    }
    t_idle = 0;
    // ---- consume slot `tail`
    mbar_wait(&bars[tail], (par >> tail) & 1u);
    par ^= 1u << tail;
    flush_signal();
    const int4 mt = meta[tail];
    unsigned char* slot = ring + (size_t)tail * cfg.SLOTB;
    const int type = mt.x & 3, t = mt.y, gi = mt.z;
    if (mt.x & kMetaFirst) {
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
      cons.begin_task(gi, slot + cfg.RSg + aux_row_bytes);
      task_n = 0;
    }
    ++task_n;
    if (type == kMetaA) {
      cons.stage_a(t, slot, aux_row(gi, t));
      if (lane == 0) bulk_commit();  // one (empty) group per consumed slot keeps the wait below uniform
      if (mt.x & kMetaLast) {
        // the release of this task is issued when the next slab has landed (or the warp runs dry): the emission stores
        // have reached L2 by then and the fence does not wait for them
        sig_ctr = cfg.doneA + gi;
        sig_n = task_n;
      }
    } else {
      cons.stage_c(t, slot, reinterpret_cast<const float*>(slot + cfg.RSg));
      fence_proxy_async();  // the slab is read by the async proxy (bulk store)
      __syncwarp();
      if (lane == 0) cons.issue_store(slot, t, pol_stream);
      // (relaxed: the aux rows of this task have been read -- they sit in shared memory -- and nothing was written
      // that another thread reads)
      if ((mt.x & kMetaLast) && lane == 0) atomicAdd(cfg.doneC + gi, task_n);
    }
    tail = tail + 1 == D ? 0 : tail + 1;
    --inflight;
    // the slot refilled next was consumed one iteration ago: its bulk store may only be the older of two pending
    if (lane == 0) bulk_wait_read<1>();
    __syncwarp();
  }
  if (lane == 0) bulk_wait_all();
}

template <int NS, int LPR>
__device__ __forceinline__ void chain_warp_main(const Problem& P, const PipeCfg& cfg, int lane, uint32_t ring) {
  constexpr int GB = 32 / LPR, Lpad = 16 * NS;
  const int T = (int)P.T;
  for (;;) {
    int q = 0;
    if (lane == 0) q = atomicAdd(cfg.ctr + 1, 1);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= (int)P.B) break;
    const int b = q, gi = b / GB, sq = b - gi * GB;
    const int4 h = __ldg(&cfg.hdr[b]);
    if (h.x > 0) {
      // all stage-A tasks of the group
      unsigned long long t0 = 0;
      for (;;) {
        int ok = 0;
        if (lane == 0) ok = ld_acquire(cfg.doneA + gi) >= __ldg(&cfg.grp[gi].x);
        if (__shfl_sync(0xffffffffu, ok, 0)) break;
        const unsigned long long now = gtime();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000ull) __trap();  // (no call here: a callee shared with the row warps would
                                                      // pin the chain warps to the row warps' register budget)
        __nanosleep(256);
      }
      ChainArgs a;
      const size_t slot = (size_t)(gi % cfg.NGS);
      const size_t TP = (size_t)T + 2 * kPadRows;
      a.paux = cfg.aux + (slot * TP + kPadRows) * cfg.AUXF + sq * Lpad;
      a.pstride = cfg.AUXF;
      a.ab = cfg.ab + ((slot * GB + sq) * TP + kPadRows) * (size_t)(16 * CRing<NS>::NSP);
      a.ring = ring;
      a.zeros = cfg.zeros;
      a.Tb = h.x; a.Lb = h.y; a.wgt = __int_as_float(h.w);
      a.want_grad = cfg.want_grad;
      a.loss_out = P.loss + b;
      chain_sequence<NS>(a, lane);
    }
    __syncwarp();
    if (lane == 0) red_release_add(cfg.doneB + gi, 1);
  }
}

template <int NS, int LPR, int CPL>
__global__ void __launch_bounds__(PipeTraits<NS>::kThreads, 1) nbctc_pipe_kernel(const Problem P, const PipeCfg cfg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + cfg.o_bar);
  for (int i = tid; i < cfg.NRW * cfg.D; i += blockDim.x) mbar_init(&bars[i], 1);
  if (tid < 4) reinterpret_cast<int*>(smem_raw + cfg.o_flags)[tid] = 0;
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  // register budget per role (NRW is a multiple of 4: the roles are whole warpgroups)
  // (the block always has kRowWarps row warps; those beyond cfg.NRW have no ring and leave at once)
  if (warp < PipeTraits<NS>::kRowWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PipeTraits<NS>::kRowRegs));
    if ((cfg.phase_mask & 5) && warp < cfg.NRW) row_warp_main<NS, LPR, CPL>(P, cfg, smem_raw, warp, lane);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PipeTraits<NS>::kChainRegs));
    if (cfg.phase_mask & 2)
      chain_warp_main<NS, LPR>(P, cfg, lane, smem_u32(smem_raw + cfg.o_cring) + (warp - PipeTraits<NS>::kRowWarps) * CRing<NS>::BYTES);
  }
}

template <int NS, int LPR, int CPL>
int launch_pipe_inst(const Problem& p, const PipeCfg& cfg, cudaStream_t stream) {
  auto kern = nbctc_pipe_kernel<NS, LPR, CPL>;
  NBCTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes));
  kern<<<cfg.grid, PipeTraits<NS>::kThreads, cfg.smem_bytes, stream>>>(p, cfg);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

template <int NS>
int launch_pipe_ns(const Problem& p, const PipeCfg& cfg, cudaStream_t stream) {
#define NBCTC_PIPE_CASE(L, Cp) \
  if (cfg.LPR == L && cfg.CPL == Cp) return launch_pipe_inst<NS, L, Cp>(p, cfg, stream);
#ifdef NBCTC_PIPE_FEW  // development builds: the instances of the BASELINE shapes only
  if constexpr (NS <= 4) { NBCTC_PIPE_CASE(8, 5) }
  if constexpr (NS == 16) { NBCTC_PIPE_CASE(32, 8) }
#else
  // LPR >= NS keeps the emission gather / gamma scatter at <= 16 states per lane
  if constexpr (NS <= 4) {
    NBCTC_PIPE_CASE(4, 1) NBCTC_PIPE_CASE(4, 2) NBCTC_PIPE_CASE(4, 3) NBCTC_PIPE_CASE(4, 4)
  }
  if constexpr (NS <= 8) {
    NBCTC_PIPE_CASE(8, 2) NBCTC_PIPE_CASE(8, 3) NBCTC_PIPE_CASE(8, 4) NBCTC_PIPE_CASE(8, 5) NBCTC_PIPE_CASE(8, 6)
    NBCTC_PIPE_CASE(8, 7) NBCTC_PIPE_CASE(8, 8)
  }
  NBCTC_PIPE_CASE(16, 1) NBCTC_PIPE_CASE(16, 2) NBCTC_PIPE_CASE(16, 3) NBCTC_PIPE_CASE(16, 4)
  NBCTC_PIPE_CASE(32, 1) NBCTC_PIPE_CASE(32, 2) NBCTC_PIPE_CASE(32, 3) NBCTC_PIPE_CASE(32, 4) NBCTC_PIPE_CASE(32, 6)
  NBCTC_PIPE_CASE(32, 8)
#endif
#undef NBCTC_PIPE_CASE
  set_error("no pipeline kernel instance for NS=%d LPR=%d CPL=%d", NS, cfg.LPR, cfg.CPL);
  return NBCTC_ERR_UNSUPPORTED;
}

}  // namespace pipe
#endif  // __CUDACC__

}  // namespace nbctc
