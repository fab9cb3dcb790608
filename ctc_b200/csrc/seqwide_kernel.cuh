// Sequence-per-warp kernel for WIDE rows (C a multiple of 4, C <= 1024, Lmax <= 256; host side in nbctc_seqwarp.cu):
// the long-sequence shape of BASELINE.json configs[3] (T=4096, C=1024, L<=256), where a (t,b) row is 4 KB and the
// gradient rows of a sequence (16 MB) cannot wait in L2 for the backward pass whatever the schedule.
//
// Same algorithm as seqwarp_kernel.cuh (one warp = one sequence, phase 1 alpha upwards, phase 2 alpha replay + beta +
// gradient downwards, per-lane power-of-two scales, 3 passes over the logits' bytes).  What differs is how rows move:
//   * every warp owns a ring of D row slots in shared memory.  Lane 0 requests whole rows with TMA bulk copies
//     (cp.async.bulk global -> shared, one mbarrier per slot) D-1 rows ahead of their use: no registers and no issue
//     slots of the other lanes are spent on latency, and 7 warps x D rows x 4 KB keep > 100 KB per SM in flight;
//   * a lane reads the row as float4 (conflict-free), the per-state label gathers come from the same slot;
//   * phase 2 turns the slot into the gradient row IN PLACE (w softmax, then the first state of every label overwrites
//     its class with the corrected value) and lane 0 sends it to the gradient tensor with ONE bulk store; the slot is
//     refilled one step later, when that store has read it (cp.async.bulk.wait_group.read).  Seven slots: the tile of 4
//     rows being worked on + the first 3 rows of the next tile, requested lowest row first (the order the alpha replay
//     consumes them).
// Template parameters: NS states per lane (Lmax <= 32 NS), NV float4 chunks per lane (C <= 128 NV).
#pragma once

#include "seqwarp_kernel.cuh"
#include "stream_kernel.cuh"  // mbarrier / bulk-copy PTX helpers

namespace nbctc {

struct WideParams {
  Problem p;
  float* lse2;       // [B][Tp]
  double* ckx;       // [B][K][32 NS]
  int* cke;          // [B][K/2 + 1][32]
  const int* order;  // null: identity
  int* ticket;       // null: one sequence per warp
  int K, Tp;
  int D;             // ring slots per warp
  int* floor_flag;   // [B] (seqwarp_kernel.cuh: SwParams::floor_flag)
  uint32_t o_gam, o_nxt, o_ring, smem_bytes;
};

template <int NV>
int launch_seqwide_nv(const WideParams& P, int NS, int grid, cudaStream_t stream);

#ifdef __CUDACC__
namespace swd {

using namespace sw;
using stream::bulk_commit;
using stream::bulk_g2s_hint;
using stream::bulk_s2g_hint;
using stream::bulk_wait_all;
using stream::bulk_wait_read;
using stream::fence_proxy_async;
using stream::mbar_arrive_expect_tx;
using stream::mbar_init;
using stream::mbar_wait;
using stream::policy_evict_first;
using stream::smem_u32;

constexpr int kMaxD = 8;

template <int NS, int NV, bool FULL>
struct Wide {
  static constexpr int Lpad = 32 * NS;
  const WideParams& P;
  const int lane;
  uint64_t* bar;        // [D]
  float* gam;           // [Lpad + 1]
  unsigned short* nxt;  // [Lpad + 1]
  unsigned char* ring;  // [D][RB]
  int C, C4, RB, D;
  uint32_t par;         // phase parity of every slot's mbarrier (kept across sequences)
  uint64_t pol;

  __device__ __forceinline__ Wide(const WideParams& P_, int lane_, unsigned char* smem) : P(P_), lane(lane_) {
    bar = reinterpret_cast<uint64_t*>(smem);
    gam = reinterpret_cast<float*>(smem + P.o_gam);
    nxt = reinterpret_cast<unsigned short*>(smem + P.o_nxt);
    ring = smem + P.o_ring;
    C = (int)P.p.C;
    C4 = C / 4;
    RB = C * 4;
    D = P.D;
    par = 0;
    pol = policy_evict_first();
  }
  __device__ __forceinline__ float4* row4(int slot) const { return reinterpret_cast<float4*>(ring + (size_t)slot * RB); }
  __device__ __forceinline__ float* rowf(int slot) const { return reinterpret_cast<float*>(ring + (size_t)slot * RB); }
  __device__ __forceinline__ void issue_load(int slot, const float* src) const {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bar[slot], (uint32_t)RB);
      bulk_g2s_hint(smem_u32(ring + (size_t)slot * RB), reinterpret_cast<uint64_t>(src), (uint32_t)RB, smem_u32(&bar[slot]), pol);
    }
  }
  __device__ __forceinline__ void wait_load(int slot) {
    mbar_wait(&bar[slot], (par >> slot) & 1u);
    par ^= 1u << slot;
  }
  __device__ __forceinline__ void load_chunks(int slot, float4 (&v)[NV]) const {
    const float4* r = row4(slot);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      v[i] = (FULL || c4 < C4) ? r[c4] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
  }
  __device__ __forceinline__ void zero_rows(int b, int t0, int t1) const {
    float4* g = reinterpret_cast<float4*>(P.p.grad + ((int64_t)t0 * P.p.B + b) * C);
    const int64_t s4 = P.p.B * (int64_t)C4;
    for (int t = t0; t < t1; ++t, g += s4)
      for (int c4 = lane; c4 < C4; c4 += 32) g[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void emissions(int slot, float nl, const int (&lab)[NS], uint32_t actm, float (&pe)[NS]) const {
    const float* r = rowf(slot);
#pragma unroll
    for (int j = 0; j < NS; ++j) pe[j] = (actm >> j) & 1u ? fmaxf(ex2f(fmaf(r[lab[j]], kL2E, nl)), kPFloor) : 0.f;
  }
  __device__ __forceinline__ void emissions(int slot, float nl, const int (&lab)[NS], uint32_t actm, float (&pe)[NS], bool& low) const {
    const float* r = rowf(slot);
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const float raw = ex2f(fmaf(r[lab[j]], kL2E, nl));
      const bool a = (actm >> j) & 1u;
      low |= a && raw < kPFloor;
      pe[j] = a ? fmaxf(raw, kPFloor) : 0.f;
    }
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
    __syncwarp();
  }

  __device__ void run(int b);
};

template <int NS, int NV, bool FULL>
__device__ void Wide<NS, NV, FULL>::run(int b) {
  const Problem& p = P.p;
  const int T = (int)p.T, B = (int)p.B;
  const int64_t Tb64 = p.in_len[b], Lb64 = p.tgt_len[b];
  bool feas = seq_feasible(Tb64, Lb64, p.T, p.Lmax) && Lb64 <= Lpad;
  // (warp reductions of warp-uniform values: the compiler then KNOWS they are uniform, keeps them in uniform registers
  // and drops the divergence guards around every shuffle inside the loops they bound)
  const int Tb = __reduce_max_sync(kFull, feas ? (int)Tb64 : 0), Lb = __reduce_max_sync(kFull, feas ? (int)Lb64 : 0);
  int lab[NS];
  uint32_t actm = 0;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int s = lane * NS + j;
    const bool a = s < Lb;
    lab[j] = a ? p.labels[(int64_t)b * p.Lmax + s] : 0;
    if (a && (lab[j] < 0 || lab[j] >= C)) feas = false;
    actm |= a ? 1u << j : 0u;
  }
  feas = __all_sync(kFull, feas);
  const float w = p.w_scalar * (p.seq_w ? p.seq_w[b] : 1.f);
  if (lane == 0) P.floor_flag[b] = 0;
  if (!feas) {
    if (lane == 0) p.loss[b] = INFINITY;
    if (p.grad) zero_rows(b, 0, T);
    return;
  }
  uint32_t leadm = 0;
  int nx1[NS];
  int R = 0;
  {
    int* labs = reinterpret_cast<int*>(gam);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NS; ++j) labs[lane * NS + j] = (actm >> j) & 1u ? lab[j] : -1 - (lane * NS + j);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int s = lane * NS + j;
      int rank = 0, nx = Lpad;
      if ((actm >> j) & 1u) {
        for (int s2 = 0; s2 < s; ++s2) rank += labs[s2] == lab[j];
        for (int s2 = Lb - 1; s2 > s; --s2) nx = labs[s2] == lab[j] ? s2 : nx;
        if (rank == 0) leadm |= 1u << j;
      }
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
  const int64_t strideT = p.B * p.C;
  const float* const seq0 = p.logits + (int64_t)b * C;  // row t = 0
  float* const lse_ws = P.lse2 + (int64_t)b * P.Tp;
  double* const ckx = P.ckx + ((int64_t)b * P.K * 32 + lane) * NS;
  int* const cke = P.cke + (int64_t)b * (P.K / 2 + 1) * 32 + lane;
  // only the lanes that hold states take part in the checkpoint traffic
  const bool ck_lane = actm != 0;

  // ================================================================ phase 1
  double x[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) x[j] = ((lane * NS + j) & 1) ? -1.0 : 1.0;
  int e = 0;
  double fac = lane == 0 ? 0.0 : 1.0;
  bool low = false;  // an emission below the float32 floor
  {
    const int npre = min(D, Tb);
    for (int t = 0; t < npre; ++t) issue_load(t, seq0 + (int64_t)t * strideT);
    int slot = 0;
    for (int t = 0; t < Tb; ++t) {
      wait_load(slot);
      float4 v[NV];
      load_chunks(slot, v);
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < NV; ++i) m = fmaxf(fmaxf(fmaxf(m, v[i].x), fmaxf(v[i].y, v[i].z)), v[i].w);
      const float nm = -kL2E * redux_max(m);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        s += (ex2f(fmaf(v[i].x, kL2E, nm)) + ex2f(fmaf(v[i].y, kL2E, nm))) + (ex2f(fmaf(v[i].z, kL2E, nm)) + ex2f(fmaf(v[i].w, kL2E, nm)));
      const float nl = nm - lg2f(warp_sum1(s));
      float pe[NS];
      emissions(slot, nl, lab, actm, pe, low);
      __syncwarp();  // every lane is done with the slot
      if (t + D < Tb) issue_load(slot, seq0 + (int64_t)(t + D) * strideT);
      if (want_grad && lane == 0) lse_ws[t] = -nl;
      if ((t & 3) == 0 && t > 0) {
        const int k = t >> 2;
        if ((k & 1) == 0) {
          lane_rescale<NS, true>(x, e, fac, lane);
          if (want_grad) cke[(k >> 1) * 32] = e;
        }
        if (want_grad && ck_lane) {
#pragma unroll
          for (int j = 0; j < NS; j += 2) {
            if (NS == 1) ckx[(int64_t)k * Lpad] = x[0];
            else *reinterpret_cast<double2*>(ckx + (int64_t)k * Lpad + j) = make_double2(x[j], x[(j + 1) % NS]);
          }
        }
      }
      alpha_step<NS>(x, pe, fac);
      if (++slot == D) slot = 0;
    }
  }
  if (__any_sync(kFull, low)) {  // the log-domain repair kernel redoes this sequence
    if (lane == 0) P.floor_flag[b] = 1;
    return;
  }
  // ---- read-out
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
    if (ok) {
      zhat *= pow2z(1023 - ezf);
      Ez += ezf - 1023;
    }
    if (lane == 0) p.loss[b] = ok ? (float)(-(log(zhat) + (double)Ez * 0.6931471805599453)) : INFINITY;
    zinv = ok ? (double)w / zhat : 0.0;
    if (!ok) {
      if (p.grad) zero_rows(b, 0, T);
      return;
    }
  }
  if (!p.grad) return;
  if (!want_grad) {
    zero_rows(b, 0, T);
    return;
  }
  __syncwarp();

  // ================================================================ phase 2
  const float lw = lg2f(w);
  double u[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int s = lane * NS + j;
    u[j] = s < Lb ? (((Lb - 1 - s) & 1) ? -1.0 : 1.0) : 0.0;
  }
  int eb = 0;
  double facb = lane == 31 ? 0.0 : 1.0;
  float* const gseq0 = p.grad + (int64_t)b * C;
  const int Kf = Tb / kTT, nrem = Tb - Kf * kTT;

  // alpha state, lane scale in front of tile k; then the scale factors of gamma (seqwarp_kernel.cuh)
  auto load_ck = [&](int k, double (&xa)[NS], double& ga, double& gb, double& faca) {
    int ea = 0;
    if (k == 0) {
#pragma unroll
      for (int j = 0; j < NS; ++j) xa[j] = ((lane * NS + j) & 1) ? -1.0 : 1.0;
    } else {
      if (ck_lane) {
#pragma unroll
        for (int j = 0; j < NS; j += 2) {
          if (NS == 1) {
            xa[0] = ckx[(int64_t)k * Lpad];
          } else {
            const double2 t2 = *reinterpret_cast<const double2*>(ckx + (int64_t)k * Lpad + j);
            xa[j] = t2.x;
            xa[(j + 1) % NS] = t2.y;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < NS; ++j) xa[j] = 0.0;
      }
      if (k >= 2) ea = cke[(k >> 1) * 32];
    }
    const int H = ea + eb - Ez;
    const int Ha = max(min(H, 0), -1000);
    ga = pow2z(Ha);  // applied to the replayed alpha AFTER the recursion (seqwarp_kernel.cuh: gscales)
    gb = pow2z(H - Ha) * zinv;
    const int eu = __shfl_up_sync(kFull, ea, 1);
    faca = lane == 0 ? 0.0 : pow2z(eu - ea);
  };
  // one backward step on the row in `slot` (time t): beta, gamma, the slot becomes the gradient row in place and
  // leaves with one bulk store
  auto grad_step = [&](int slot, int t, const double (&a)[NS], const float (&pe)[NS], float nl, double gb) {
    double bt[NS];
    beta_step<NS>(u, bt, pe, facb);
    float g[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) g[j] = (float)(a[j] * clamp_big(bt[j] * gb));
    const float nlw = nl + lw;
    float4* r4 = row4(slot);
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const int c4 = lane + 32 * q;
      if (FULL || c4 < C4) {
        float4 v = r4[c4];
        v.x = ex2f(fmaf(v.x, kL2E, nlw));
        v.y = ex2f(fmaf(v.y, kL2E, nlw));
        v.z = ex2f(fmaf(v.z, kL2E, nlw));
        v.w = ex2f(fmaf(v.w, kL2E, nlw));
        r4[c4] = v;
      }
    }
    combine(g, nx1, R);  // (its barrier also orders the row writes before the corrected entries)
    float* rf = rowf(slot);
#pragma unroll
    for (int j = 0; j < NS; ++j)
      if ((leadm >> j) & 1u) rf[lab[j]] = fmaf(w, pe[j], -g[j]);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      bulk_s2g_hint(reinterpret_cast<uint64_t>(gseq0 + (int64_t)t * strideT), smem_u32(r4), (uint32_t)RB, pol);
      bulk_commit();
    }
  };

  if (nrem > 0) {  // the sequence's last steps, one at a time through slot 0
    const int k = Kf;
    double xa0[NS], ga, gb, faca;
    load_ck(k, xa0, ga, gb, faca);
    for (int i = nrem - 1; i >= 0; --i) {
      const int t = k * kTT + i;
      issue_load(0, seq0 + (int64_t)t * strideT);
      double xa[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) xa[j] = xa0[j];
      float pr[NS], nl = 0.f;
      for (int ii = 0; ii <= i; ++ii) {  // replay up to step i (label entries straight from global memory)
        nl = -lse_ws[k * kTT + ii];
        const float* rp = seq0 + (int64_t)(k * kTT + ii) * strideT;
#pragma unroll
        for (int j = 0; j < NS; ++j) pr[j] = (actm >> j) & 1u ? fmaxf(ex2f(fmaf(__ldg(rp + lab[j]), kL2E, nl)), kPFloor) : 0.f;
        alpha_step<NS>(xa, pr, faca);
      }
      wait_load(0);
#pragma unroll
      for (int j = 0; j < NS; ++j) xa[j] *= ga;
      grad_step(0, t, xa, pr, nl, gb);
      if (lane == 0) bulk_wait_all();
      __syncwarp();
    }
    lane_rescale<NS, false>(u, eb, facb, lane);
  }

  if (Kf > 0) {
    // Seven slots: the four rows of the tile being worked on and the first three of the next tile down.  Rows are
    // requested in the order the alpha replay needs them (a tile's lowest row first), so the row a tile waits for
    // first has been in flight for a whole tile.  A slot is refilled one step after its bulk store was issued.
    int cur[kTT] = {0, 1, 2, 3};  // slot of row t0 + i
    int nx3[3] = {4, 5, 6};       // slot of row t0 - 4 + r
    {
      const int t0 = (Kf - 1) * kTT;
#pragma unroll
      for (int i = 0; i < kTT; ++i) issue_load(cur[i], seq0 + (int64_t)(t0 + i) * strideT);
#pragma unroll
      for (int r = 0; r < 3; ++r)
        if (t0 - kTT + r >= 0) issue_load(nx3[r], seq0 + (int64_t)(t0 - kTT + r) * strideT);
    }
    int pend_slot = -1, pend_t = -1;  // slot whose store was issued in the previous step, and the row it takes next
    int since = 0;
    for (int k = Kf - 1; k >= 0; --k) {
      const int t0 = k * kTT;
      double xa[NS], ga, gb, faca;
      load_ck(k, xa, ga, gb, faca);
      if (k > 0) {  // the next tile's checkpoint (lanes 0 .. 2 NS - 1: its 128-byte lines), lane scales and log-partitions
        const char* a = lane < 2 * NS ? reinterpret_cast<const char*>(ckx - lane * NS + (int64_t)(k - 1) * Lpad) + 128 * lane
                      : lane == 2 * NS ? reinterpret_cast<const char*>(cke - lane + ((k - 1) >> 1) * 32)
                                       : reinterpret_cast<const char*>(lse_ws + (k - 1) * kTT);
        if (lane <= 2 * NS + 1) pf_line_l1(a);
      }
      const float4 l4 = *reinterpret_cast<const float4*>(lse_ws + t0);
      const float nl[kTT] = {-l4.x, -l4.y, -l4.z, -l4.w};
      float pe[kTT][NS];
      double a[kTT][NS];
#pragma unroll
      for (int i = 0; i < kTT; ++i) {
        wait_load(cur[i]);
        emissions(cur[i], nl[i], lab, actm, pe[i]);
        alpha_step<NS>(xa, pe[i], faca);
#pragma unroll
        for (int j = 0; j < NS; ++j) a[i][j] = xa[j] * ga;
      }
#pragma unroll
      for (int i = kTT - 1; i >= 0; --i) {
        grad_step(cur[i], t0 + i, a[i], pe[i], nl[i], gb);
        if (pend_slot >= 0 && pend_t >= 0) {
          if (lane == 0) bulk_wait_read<1>();  // the previous step's store has read its slot
          issue_load(pend_slot, seq0 + (int64_t)pend_t * strideT);
        }
        pend_slot = cur[i];
        pend_t = i == kTT - 1 ? t0 - 1 : t0 - 2 * kTT + (2 - i);  // (k-1, row 3), then (k-2, rows 0, 1, 2)
      }
      // the next tile down: its rows 0..2 are in nx3, row 3 takes the slot of this tile's row 3
      const int c0 = cur[0], c1 = cur[1], c2 = cur[2];
      cur[0] = nx3[0]; cur[1] = nx3[1]; cur[2] = nx3[2];
      nx3[0] = c2; nx3[1] = c1; nx3[2] = c0;
      if (++since == 2) {
        since = 0;
        lane_rescale<NS, false>(u, eb, facb, lane);
      }
    }
  }
  if (lane == 0) bulk_wait_all();  // the ring is reused by the next sequence; the stores must have left shared memory
  __syncwarp();
  if (Tb < T) zero_rows(b, Tb, T);
}

template <int NS, int NV, bool FULL>
__global__ void __launch_bounds__(32, NS >= 8 ? 7 : 8) seqwide_kernel(const __grid_constant__ WideParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x;
  Wide<NS, NV, FULL> wd(P, lane, smem);
  if (lane == 0) {
    for (int i = 0; i < P.D; ++i) mbar_init(&wd.bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int B = (int)P.p.B;
  int i = blockIdx.x;
  for (;;) {
    if (P.ticket) {
      if (lane == 0) i = atomicAdd(P.ticket, 1);
      i = __shfl_sync(kFull, i, 0);
    }
    if (i >= B) break;
    wd.run(P.order ? P.order[i] : i);
    i += gridDim.x;
  }
}

template <int NS, int NV, bool FULL>
int launch_wide_full(const WideParams& P, int grid, cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    NBCTC_CUDA_CHECK(cudaFuncSetAttribute(seqwide_kernel<NS, NV, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 31 * 1024));
    attr_done = true;
  }
  seqwide_kernel<NS, NV, FULL><<<grid, 32, P.smem_bytes, stream>>>(P);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}
template <int NS, int NV>
int launch_wide_one(const WideParams& P, int grid, cudaStream_t stream) {
  // rows that fill every lane's chunks (C = 128 NV) run without the per-chunk bounds predicates
  return P.p.C == 128 * NV ? launch_wide_full<NS, NV, true>(P, grid, stream) : launch_wide_full<NS, NV, false>(P, grid, stream);
}

}  // namespace swd

template <int NV>
int launch_seqwide_nv(const WideParams& P, int NS, int grid, cudaStream_t stream) {
  switch (NS) {
    case 1: return swd::launch_wide_one<1, NV>(P, grid, stream);
    case 2: return swd::launch_wide_one<2, NV>(P, grid, stream);
    case 4: return swd::launch_wide_one<4, NV>(P, grid, stream);
    case 8: return swd::launch_wide_one<8, NV>(P, grid, stream);
    default: set_error("seqwide: NS=%d not built", NS); return NBCTC_ERR_UNSUPPORTED;
  }
}
#endif  // __CUDACC__

}  // namespace nbctc
