// Host side of the sequence-per-warp kernel (seqwarp_kernel.cuh): shape gate, workspace carve-up, the longest-first
// work queue for batches larger than one wave (the north star's "length-bucketed launch fusion": ONE launch, the
// sequences sorted into 256 length buckets on the device and handed to the persistent warps longest first), dispatch.
//
// Algorithmic HBM bytes per launch: 2*4*T*B*C (logits read once, gradient written once).  The kernel itself moves
// 3*4*T*B*C (the logits are read again in phase 2) + 4 bytes per row of log-partitions + one 8-byte alpha
// checkpoint per state and 4 steps, both ways.
#include <stdlib.h>

#include <algorithm>

#include "seqwide_kernel.cuh"

namespace nbctc {
namespace {

constexpr int kPrepThreads = 1024;
constexpr int kBins = 256;

// order[] = sequences by descending input length (256 buckets; ascending b inside a bucket up to the scheduling of
// one chunk of 1024 threads -- the order only steers the work queue, never the results); ticket = 0
__global__ void __launch_bounds__(kPrepThreads) seqwarp_prep_kernel(const int64_t* __restrict__ in_len, int B, int T, int* __restrict__ order,
                                                                     int* __restrict__ ticket) {
  __shared__ int hist[kBins];
  __shared__ int cursor[kBins];
  for (int i = threadIdx.x; i < kBins; i += kPrepThreads) hist[i] = 0;
  __syncthreads();
  auto bin_of = [&](int b) {
    int64_t tb = in_len[b];
    tb = tb < 0 ? 0 : tb > T ? T : tb;
    return (int)(((int64_t)(T - tb) * (kBins - 1)) / (T > 0 ? T : 1));
  };
  for (int b = threadIdx.x; b < B; b += kPrepThreads) atomicAdd(&hist[bin_of(b)], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < kBins; ++i) {
      cursor[i] = acc;
      acc += hist[i];
    }
    *ticket = 0;
  }
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += kPrepThreads) {
    const int b = b0 + threadIdx.x;
    if (b < B) order[atomicAdd(&cursor[bin_of(b)], 1)] = b;
    __syncthreads();
  }
}

struct SwPlan {
  bool ok;
  bool wide;        // seqwide_kernel.cuh (rows through a shared-memory ring) instead of seqwarp_kernel.cuh (registers)
  int NS, EPL, K, Tp;
  int NV, D;        // wide: float4 chunks per lane, ring slots per warp
  uint32_t o_gam, o_nxt, o_ring, smem_bytes;
  size_t o_order, o_flag, o_lse, o_ckx, o_cke, bytes;
};

constexpr size_t kWideSmemBudget = 31 * 1024;  // per warp-CTA: 7 of them (+1 KB each of driver reserve) share an SM

SwPlan make_plan(int64_t T, int64_t B, int64_t C, int64_t Lmax) {
  SwPlan pl{};
  if (T < 1 || B < 1 || C < 1 || Lmax < 1 || T > (1 << 24) || B > (1 << 24) || B * C >= ((int64_t)1 << 26)) return pl;
  if (T * B * C >= ((int64_t)1 << 40)) return pl;
  if (Lmax <= 64 && C <= 256) {
    pl.NS = Lmax <= 32 ? 1 : 2;
    pl.EPL = (int)((C + 31) / 32);
  } else if (Lmax <= 256 && C <= 1024 && C % 4 == 0 && C >= 16) {
    pl.wide = true;
    pl.NS = Lmax <= 32 ? 1 : Lmax <= 64 ? 2 : Lmax <= 128 ? 4 : 8;
    pl.NV = C <= 512 ? 4 : 8;
    size_t off = 64;  // mbarriers
    pl.o_gam = (uint32_t)off;
    off = align_up(off + sizeof(float) * (32 * pl.NS + 4), 16);
    pl.o_nxt = (uint32_t)off;
    off = align_up(off + sizeof(unsigned short) * (32 * pl.NS + 4), 128);
    pl.o_ring = (uint32_t)off;
    const int64_t d = ((int64_t)kWideSmemBudget - (int64_t)off) / (C * 4);
    if (d < 7) return pl;  // a tile of 4 rows + the first 3 of the next one
    pl.D = 7;
    pl.smem_bytes = (uint32_t)(off + (size_t)pl.D * C * 4);
  } else {
    return pl;
  }
  pl.K = (int)((T + 3) / 4);
  pl.Tp = pl.K * 4;
  size_t off = 256;
  pl.o_order = off;
  off = align_up(off + sizeof(int) * (size_t)B, 256);
  pl.o_flag = off;  // [B] sequences for the log-domain repair kernel
  off = align_up(off + sizeof(int) * (size_t)B, 256);
  pl.o_lse = off;
  if (pl.wide) {
    off = align_up(off + sizeof(float) * (size_t)B * pl.Tp, 256);
    pl.o_ckx = off;
    off = align_up(off + sizeof(double) * (size_t)B * pl.K * 32 * pl.NS, 256);
    pl.o_cke = off;
    off = align_up(off + sizeof(int) * (size_t)B * (pl.K / 2 + 1) * 32, 256);
  } else {  // one record per sequence (seqwarp_launch)
    const size_t rec = align_up(align_up(align_up(sizeof(float) * (size_t)pl.Tp, 128) + sizeof(int) * (size_t)(pl.K / 2 + 1) * 32, 256) +
                                    sizeof(double) * (size_t)pl.K * 32 * pl.NS, 256);
    if (rec >= ((size_t)1 << 30)) return pl;  // 32-bit offsets inside a record
    off = align_up(off + rec * (size_t)B, 256);
  }
  pl.bytes = off;
  pl.ok = true;
  return pl;
}

int g_sms = 0;
int g_blocks_per_sm[3] = {0, 0, 0};

}  // namespace

bool seqwarp_supported(int64_t T, int64_t B, int64_t C, int64_t Lmax) { return make_plan(T, B, C, Lmax).ok; }
bool seqwarp_is_wide(int64_t T, int64_t B, int64_t C, int64_t Lmax) { return make_plan(T, B, C, Lmax).wide; }

size_t seqwarp_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax) {
  const SwPlan pl = make_plan(T, B, C, Lmax);
  return pl.ok ? pl.bytes : 256;
}

int seqwarp_launch(const Problem& p, void* ws, size_t ws_bytes, const float* row_lse_in, float* row_lse_out, cudaStream_t stream) {
  const SwPlan pl = make_plan(p.T, p.B, p.C, p.Lmax);
  if (!pl.ok) {
    set_error("shape not supported by the sequence-per-warp kernel");
    return NBCTC_ERR_UNSUPPORTED;
  }
  if (ws == nullptr || ws_bytes < pl.bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", pl.bytes, ws_bytes);
    return NBCTC_ERR_WORKSPACE;
  }
  if (g_sms == 0) {
    int dev = 0;
    NBCTC_CUDA_CHECK(cudaGetDevice(&dev));
    NBCTC_CUDA_CHECK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  char* w = static_cast<char*>(ws);
  if (pl.wide) {
    if (row_lse_in || row_lse_out) {
      set_error("row_lse is not available on the wide-row kernel (C > 256 or Lmax > 64)");
      return NBCTC_ERR_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(p.logits) & 15) != 0 || (reinterpret_cast<uintptr_t>(p.grad) & 15) != 0) {
      set_error("the wide-row kernel needs 16-byte aligned logits / grad_logits");
      return NBCTC_ERR_UNSUPPORTED;
    }
    WideParams W{};
    W.p = p;
    W.lse2 = reinterpret_cast<float*>(w + pl.o_lse);
    W.ckx = reinterpret_cast<double*>(w + pl.o_ckx);
    W.cke = reinterpret_cast<int*>(w + pl.o_cke);
    W.K = pl.K;
    W.Tp = pl.Tp;
    W.D = pl.D;
    W.floor_flag = reinterpret_cast<int*>(w + pl.o_flag);
    W.o_gam = pl.o_gam; W.o_nxt = pl.o_nxt; W.o_ring = pl.o_ring; W.smem_bytes = pl.smem_bytes;
    const int64_t resident = (int64_t)(pl.NS >= 8 ? 7 : 8) * g_sms;
    int grid;
    if (p.B <= resident) {
      grid = (int)p.B;
    } else {
      grid = (int)resident;
      W.ticket = reinterpret_cast<int*>(w);
      W.order = reinterpret_cast<int*>(w + pl.o_order);
      seqwarp_prep_kernel<<<1, kPrepThreads, 0, stream>>>(p.in_len, (int)p.B, (int)p.T, const_cast<int*>(W.order), W.ticket);
      NBCTC_LAUNCH_CHECK();
    }
    const int rcw = pl.NV == 4 ? launch_seqwide_nv<4>(W, pl.NS, grid, stream) : launch_seqwide_nv<8>(W, pl.NS, grid, stream);
    if (rcw != NBCTC_OK) return rcw;
    LogWs lw{W.floor_flag, W.lse2, (int64_t)pl.Tp, W.ckx, (int64_t)pl.K * 32 * pl.NS};
    return logdom_repair_launch(p, lw, stream);
  }
  SwParams P{};
  P.p = p;
  // the three per-sequence arrays interleaved into one record per sequence (same total size as the separate arrays)
  P.rec = w + pl.o_lse;
  {
    const size_t lse_b = sizeof(float) * (size_t)pl.Tp, cke_b = sizeof(int) * (size_t)(pl.K / 2 + 1) * 32;
    P.o_cke = (int)align_up(lse_b, 128);
    P.o_ckx = (int)align_up(P.o_cke + cke_b, 256);
    P.rec_bytes = (int64_t)align_up(P.o_ckx + sizeof(double) * (size_t)pl.K * 32 * pl.NS, 256);
  }
  P.floor_flag = reinterpret_cast<int*>(w + pl.o_flag);
  P.row_lse_in = row_lse_in;
  P.row_lse_out = row_lse_out;
  P.K = pl.K;
  P.Tp = pl.Tp;
  const int per_sm = seqwarp_ctas_per_sm(pl.NS, pl.EPL);
  const int64_t resident = (int64_t)per_sm * g_sms;  // CTAs of one warp
  const int64_t need = p.B;
  int grid;
  if (need <= resident) {
    grid = (int)need;  // the whole batch is one wave: warp = sequence
  } else {
    grid = (int)resident;
    P.ticket = reinterpret_cast<int*>(w);
    P.order = reinterpret_cast<int*>(w + pl.o_order);
    seqwarp_prep_kernel<<<1, kPrepThreads, 0, stream>>>(p.in_len, (int)p.B, (int)p.T, const_cast<int*>(P.order), P.ticket);
    NBCTC_LAUNCH_CHECK();
  }
  const int rcs = launch_seqwarp(P, pl.NS, pl.EPL, grid, stream);
  if (rcs != NBCTC_OK) return rcs;
  // sequences with an emission below the float32 floor: redone in the log domain, in their own records
  LogWs lw{P.floor_flag, reinterpret_cast<float*>(P.rec), P.rec_bytes / 4, reinterpret_cast<double*>(P.rec + P.o_ckx), P.rec_bytes / 8};
  return logdom_repair_launch(p, lw, stream);
}

int launch_seqwarp(const SwParams& P, int NS, int EPL, int grid, cudaStream_t stream) {
  switch (EPL) {
    case 1: return launch_seqwarp_epl<1>(P, NS, grid, stream);
    case 2: return launch_seqwarp_epl<2>(P, NS, grid, stream);
    case 3: return launch_seqwarp_epl<3>(P, NS, grid, stream);
    case 4: return launch_seqwarp_epl<4>(P, NS, grid, stream);
    case 5: return launch_seqwarp_epl<5>(P, NS, grid, stream);
    case 6: return launch_seqwarp_epl<6>(P, NS, grid, stream);
    case 7: return launch_seqwarp_epl<7>(P, NS, grid, stream);
    case 8: return launch_seqwarp_epl<8>(P, NS, grid, stream);
    default: set_error("seqwarp: EPL=%d not built", EPL); return NBCTC_ERR_UNSUPPORTED;
  }
}

// resident warps (CTA = one warp) per SM: registers (28 / 20 by the launch bounds) or the row ring in shared memory
int seqwarp_ctas_per_sm(int NS, int EPL) {
  static int cache[3][9] = {};
  int& c = cache[NS][EPL];
  if (c == 0) {
    switch (EPL) {
      case 1: c = seqwarp_occupancy_epl<1>(NS); break;
      case 2: c = seqwarp_occupancy_epl<2>(NS); break;
      case 3: c = seqwarp_occupancy_epl<3>(NS); break;
      case 4: c = seqwarp_occupancy_epl<4>(NS); break;
      case 5: c = seqwarp_occupancy_epl<5>(NS); break;
      case 6: c = seqwarp_occupancy_epl<6>(NS); break;
      case 7: c = seqwarp_occupancy_epl<7>(NS); break;
      default: c = seqwarp_occupancy_epl<8>(NS); break;
    }
    if (c <= 0) c = NS == 1 ? 24 : 16;
  }
  return c;
}

}  // namespace nbctc
