// Pipeline path of the single-label loss (kernel in pipe_kernel.cuh): shape -> launch plan, workspace carve-up,
// the per-call preparation kernel (lengths, labels with duplicate ranks, group table, counters) and dispatch.
//
// HBM traffic per call: logits once (stage A) + gradient once (stage C); the second read of the logits, the
// emission / gamma tile and the stored chain states live in L2 as long as the window of groups between the two
// stages fits (cfg2: ~1 MB per group of 4 sequences); long sequences (cfg4: 16 MB of logits per sequence) re-read
// the logits from HBM: 3 passes instead of the 4 of a store-softmax-then-fix-up design.
// Algorithmic bytes per sequence: 2*4*T*C (+ labels); per real lattice cell 8*C/mean(L).
#include <stdlib.h>

#include <algorithm>

#include "pipe_kernel.cuh"

namespace nbctc {
namespace {

int env_int_cached(const char* name, int dflt, int* cache) {
  if (*cache == INT32_MIN) {
    const char* v = getenv(name);
    *cache = (v && *v) ? atoi(v) : dflt;
  }
  return *cache;
}
int g_env_tb = INT32_MIN, g_env_win = INT32_MIN, g_env_split = INT32_MIN, g_env_d = INT32_MIN, g_env_nrw = INT32_MIN,
    g_env_lpr = INT32_MIN, g_env_keep = INT32_MIN;

struct PipePlan {
  bool ok;
  PipeCfg cfg;
  size_t o_ctr, o_zeros, o_hdr, o_grp, o_lab, o_doneA, o_doneB, o_doneC, o_aux, o_ab, ws_bytes;
};

// prep kernel: one CTA per group
__global__ void __launch_bounds__(128) pipe_prep_kernel(const Problem P, const PipeCfg cfg) {
  extern __shared__ int s_lab[];  // [GB][Lpad]
  __shared__ int s_T[8], s_L[8], s_bad[8], s_rank[8];
  const int g = blockIdx.x, tid = threadIdx.x;
  const int GB = cfg.GB, Lpad = cfg.Lpad;
  const int64_t b0 = (int64_t)g * GB;
  const int gcnt = (int)min((int64_t)GB, P.B - b0);
  if (tid < 8) { s_bad[tid] = 0; s_rank[tid] = 0; }
  if (tid < GB) {
    int Tb = 0, Lb = 0;
    if (tid < gcnt) {
      const int64_t Tb64 = P.in_len[b0 + tid], Lb64 = P.tgt_len[b0 + tid];
      if (seq_feasible(Tb64, Lb64, P.T, P.Lmax)) { Tb = (int)Tb64; Lb = (int)Lb64; }
    }
    s_T[tid] = Tb; s_L[tid] = Lb;
  }
  __syncthreads();
  for (int idx = tid; idx < GB * Lpad; idx += blockDim.x) {
    const int r = idx / Lpad, s = idx - r * Lpad;
    int l = 0;
    if (r < gcnt && s < s_L[r]) {
      l = P.labels[(b0 + r) * P.Lmax + s];
      if (l < 0 || l >= P.C) { atomicOr(&s_bad[r], 1); l = 0; }
    }
    s_lab[idx] = l;
  }
  __syncthreads();
  // duplicate rank of every state among the earlier states with the same class (stage C scatters one rank per round)
  for (int idx = tid; idx < GB * Lpad; idx += blockDim.x) {
    const int r = idx / Lpad, s = idx - r * Lpad;
    int v = 0;
    if (r < gcnt && s < s_L[r] && !s_bad[r]) {
      const int l = s_lab[idx];
      int rank = 0;
      for (int q = 0; q < s; ++q) rank += (s_lab[r * Lpad + q] == l) ? 1 : 0;
      if (rank > 0) atomicMax(&s_rank[r], rank);
      v = l | (rank << stream::kLabBits);
    }
    if (r < gcnt) cfg.lab[(b0 + r) * Lpad + s] = v;
  }
  __syncthreads();
  if (tid < gcnt) {
    const bool ok = s_T[tid] > 0 && !s_bad[tid];
    const float w = P.w_scalar * (P.seq_w ? P.seq_w[b0 + tid] : 1.f);
    cfg.hdr[b0 + tid] = make_int4(ok ? s_T[tid] : 0, ok ? s_L[tid] : 0, s_rank[tid], __float_as_int(w));
    if (!ok) P.loss[b0 + tid] = INFINITY;
  }
  if (tid == 0) {
    int Tg = 0;
    for (int r = 0; r < gcnt; ++r)
      if (s_T[r] > 0 && !s_bad[r]) Tg = max(Tg, s_T[r]);
    cfg.grp[g] = make_int2(Tg, gcnt);
    cfg.doneA[g] = 0; cfg.doneB[g] = 0; cfg.doneC[g] = 0;
    if (g == 0) { cfg.ctr[0] = 0; cfg.ctr[1] = 0; }
  }
  if (g == 0 && tid < 64) cfg.zeros[tid] = 0.f;
  {
  }
}

PipePlan make_pipe_plan(int64_t T, int64_t B, int64_t C, int64_t Lmax) {
  PipePlan pl{};
  pl.ok = false;
  if (Lmax > 256 || T > ((int64_t)1 << 24) || C > 1024 + 0 || C < 4 || B > ((int64_t)1 << 28)) return pl;
  PipeCfg& c = pl.cfg;
  c.NS = Lmax <= 32 ? 2 : Lmax <= 64 ? 4 : Lmax <= 128 ? 8 : 16;
  c.Lpad = 16 * c.NS;
  // chunks per row: with C % 4 == 0 every row starts on a 16-byte boundary, otherwise at any of the 4 phases
  const int64_t nch = (C % 4 == 0) ? C / 4 : (3 + C + 3) / 4;
  const int want_lpr = env_int_cached("NBCTC_PIPE_LPR", 0, &g_env_lpr);
  int lpr = want_lpr ? want_lpr : nch <= 16 ? 4 : nch <= 64 ? 8 : nch <= 128 ? 16 : 32;
  lpr = std::max(lpr, std::max(4, c.NS));  // at most 16 states per lane in the emission gather / gamma scatter
  if (lpr != 4 && lpr != 8 && lpr != 16 && lpr != 32) return pl;
  c.LPR = lpr;
  int cpl = (int)((nch + lpr - 1) / lpr);
  if (lpr == 8) cpl = std::max(cpl, 2);
  if (lpr == 32 && cpl == 5) cpl = 6;
  if (lpr == 32 && cpl == 7) cpl = 8;
  const int cpl_max = lpr == 4 ? 4 : lpr == 8 ? 8 : lpr == 16 ? 4 : 8;
  if (cpl > cpl_max) return pl;
  c.CPL = cpl;
  c.GB = 32 / lpr;
  c.RSg = (int)align_up((size_t)c.GB * C * 4, 16) + 32;
  c.AUXF = (int)align_up((size_t)c.GB * c.Lpad + 2 * c.GB, 4);
  c.SLOTB = (int)align_up((size_t)c.RSg + (size_t)c.AUXF * 4 + (size_t)c.GB * c.Lpad * 4 + (size_t)c.GB * 16, 128);
  // ring: as many row warps and slots as fit into shared memory
  const size_t cap = 222 * 1024 - (size_t)kPipeChainWarps * pipe::pipe_chain_ring_bytes(c.NS);
  const int nrw_max = pipe_row_warps(c.NS);
  c.NRW = nrw_max;
  c.D = 4;
  const int want_d = env_int_cached("NBCTC_PIPE_D", 0, &g_env_d), want_nrw = env_int_cached("NBCTC_PIPE_NRW", 0, &g_env_nrw);
  auto fits = [&](int nrw, int d) { return (size_t)nrw * d * c.SLOTB + 4096 <= cap; };
  if (want_d && want_nrw && fits(want_nrw, want_d) && want_d >= 2 && want_d <= 4 && want_nrw <= nrw_max) {
    c.D = want_d; c.NRW = want_nrw;
  } else {
    bool placed = false;
    const int tries[][2] = {{16, 4}, {16, 3}, {12, 3}, {8, 3}, {8, 2}, {4, 2}};
    for (auto& t : tries)
      if (t[0] <= nrw_max && fits(t[0], t[1])) { c.NRW = t[0]; c.D = t[1]; placed = true; break; }
    if (!placed) return pl;
  }
  // slabs per warp and task: 1 keeps the fewest groups in flight; long sequences take more (fewer label reloads)
  // (4 when every row warp of the machine gets at least 8 tasks of that size, else fewer)
  const int64_t slabs_per_warp = ((B + c.GB - 1) / c.GB) * T / (148 * c.NRW);
  const int ks_auto = slabs_per_warp >= 32 ? 4 : slabs_per_warp >= 16 ? 2 : 1;
  const int ks = std::max(1, env_int_cached("NBCTC_PIPE_KS", ks_auto, &g_env_tb));
  c.TB = ks * c.NRW;
  c.TPG = (int)((T + c.TB - 1) / c.TB);
  c.NG = (int)((B + c.GB - 1) / c.GB);
  if ((int64_t)c.NG * c.TPG > ((int64_t)1 << 30)) return pl;
  c.grid = (int)std::max<int64_t>(1, std::min<int64_t>(148, std::max<int64_t>((int64_t)c.NG * c.TPG, (B + kPipeChainWarps - 1) / kPipeChainWarps)));
  // Window between stage A and stage C in groups: the chains that must be in flight to keep up with the rows at
  // ~60 % of the HBM roofline = sequence rate x chain latency (T steps of ~cyc cycles), with a margin.
  const double t_target_us = (2.0 * 4.0 * T * B * C) / (0.6 * 6.5e6);
  const double cyc = c.NS == 2 ? 80 : c.NS == 4 ? 110 : c.NS == 8 ? 180 : 300;
  const double latency_us = (double)T * cyc / 1900.0 + 4.0;
  int win = (int)(1.5 * (double)B / t_target_us * latency_us / c.GB) + 2;
  const int want_win = env_int_cached("NBCTC_PIPE_WIN", 0, &g_env_win);
  if (want_win > 0) win = want_win;
  win = std::max(1, std::min(win, c.NG));
  // in positions of a CTA's ticket sequence; wmax * grid >= TPG + grid keeps the schedule free of deadlocks
  c.wmax = std::max((win * c.TPG + c.grid - 1) / c.grid + 1, c.TPG / c.grid + 2);
  const bool split = env_int_cached("NBCTC_PIPE_SPLIT", 0, &g_env_split) != 0;
  // aux slots: everything stage A can be ahead of stage C, plus the groups the tickets in flight span
  const int span = c.grid / std::max(1, c.TPG) + 2;
  int64_t ngs = split ? c.NG : std::min<int64_t>(c.NG, ((int64_t)c.wmax * c.grid + c.TPG - 1) / c.TPG + 2 * span + 4);
  c.NGS = 1;
  while (c.NGS < ngs) c.NGS <<= 1;  // a power of two: slot = group & (NGS - 1)
  c.phase_mask = 7;
  c.keep_logits = env_int_cached("NBCTC_PIPE_KEEP", 2, &g_env_keep);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 16); return (uint32_t)o; };
  c.o_bar = take(sizeof(uint64_t) * c.NRW * c.D);
  c.o_meta = take(sizeof(int4) * c.NRW * c.D);
  c.o_flags = take(16);
  off = align_up(off, 128);
  c.o_ring = take((size_t)c.NRW * c.D * c.SLOTB);
  c.o_cring = take((size_t)kPipeChainWarps * pipe::pipe_chain_ring_bytes(c.NS));
  c.smem_bytes = (uint32_t)off;
  // workspace
  size_t w = 0;
  auto wtake = [&](size_t bytes) { size_t o = w; w = align_up(w + bytes, 256); return o; };
  pl.o_ctr = wtake(256);
  pl.o_zeros = wtake(256);
  pl.o_hdr = wtake(sizeof(int4) * (size_t)B);
  pl.o_grp = wtake(sizeof(int2) * (size_t)c.NG);
  pl.o_lab = wtake(sizeof(int) * (size_t)B * c.Lpad);
  pl.o_doneA = wtake(sizeof(int) * (size_t)c.NG);
  pl.o_doneB = wtake(sizeof(int) * (size_t)c.NG);
  pl.o_doneC = wtake(sizeof(int) * (size_t)c.NG);
  pl.o_aux = wtake(sizeof(float) * (size_t)c.NGS * (T + 2 * pipe::kPadRows) * c.AUXF);
  pl.o_ab = wtake(sizeof(uint32_t) * (size_t)c.NGS * c.GB * (T + 2 * pipe::kPadRows) * 16 * (c.NS + (c.NS == 2 ? 2 : 4)));
  pl.ws_bytes = w;
  pl.ok = true;
  return pl;
}

int g_sm_count = 0;

}  // namespace

bool pipe_supported(int64_t T, int64_t B, int64_t C, int64_t Lmax) { return make_pipe_plan(T, B, C, Lmax).ok; }

size_t pipe_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax) {
  PipePlan pl = make_pipe_plan(T, B, C, Lmax);
  return pl.ok ? pl.ws_bytes + 256 : 256;  // + room to align the base to 256 bytes
}

int launch_pipe_prep(const Problem& p, const PipeCfg& cfg, cudaStream_t stream) {
  pipe_prep_kernel<<<cfg.NG, 128, sizeof(int) * cfg.GB * cfg.Lpad, stream>>>(p, cfg);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

int pipe_launch(const Problem& p, void* ws, size_t ws_bytes, cudaStream_t stream) {
  PipePlan pl = make_pipe_plan(p.T, p.B, p.C, p.Lmax);
  if (!pl.ok || !fused_pointers_ok(p)) {
    set_error("shape or pointer alignment not supported by the pipeline kernel");
    return NBCTC_ERR_UNSUPPORTED;
  }
  const size_t pad = (256 - (reinterpret_cast<uintptr_t>(ws) & 255)) & 255;
  if (ws == nullptr || ws_bytes < pl.ws_bytes + pad) {
    set_error("workspace too small: need %zu bytes, got %zu", pl.ws_bytes + 256, ws_bytes);
    return NBCTC_ERR_WORKSPACE;
  }
  PipeCfg& c = pl.cfg;
  char* w = static_cast<char*>(ws) + pad;
  c.ctr = reinterpret_cast<int*>(w + pl.o_ctr);
  c.zeros = reinterpret_cast<float*>(w + pl.o_zeros);
  c.hdr = reinterpret_cast<int4*>(w + pl.o_hdr);
  c.grp = reinterpret_cast<int2*>(w + pl.o_grp);
  c.lab = reinterpret_cast<int*>(w + pl.o_lab);
  c.doneA = reinterpret_cast<int*>(w + pl.o_doneA);
  c.doneB = reinterpret_cast<int*>(w + pl.o_doneB);
  c.doneC = reinterpret_cast<int*>(w + pl.o_doneC);
  c.aux = reinterpret_cast<float*>(w + pl.o_aux);
  c.ab = reinterpret_cast<uint32_t*>(w + pl.o_ab);
  c.want_grad = p.grad != nullptr;
  if (g_sm_count == 0) {
    int dev = 0, n = 0;
    NBCTC_CUDA_CHECK(cudaGetDevice(&dev));
    NBCTC_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    g_sm_count = n;
  }
  c.grid = std::min(c.grid, g_sm_count);  // (the plan assumed 148 SMs)
  c.wmax = std::max(c.wmax, c.TPG / c.grid + 2);
  int rc = launch_pipe_prep(p, c, stream);
  if (rc != NBCTC_OK) return rc;
  auto launch = [&](const PipeCfg& cc) {
    switch (cc.NS) {
      case 2: return launch_pipe_ns2(p, cc, stream);
      case 4: return launch_pipe_ns4(p, cc, stream);
      case 8: return launch_pipe_ns8(p, cc, stream);
      default: return launch_pipe_ns16(p, cc, stream);
    }
  };
  if (env_int_cached("NBCTC_PIPE_SPLIT", 0, &g_env_split) != 0) {
    // debugging aid: the three stages as three launches (every dependency is already met when a stage starts)
    for (int ph = 0; ph < 3; ++ph) {
      if (ph == 2 && !c.want_grad) break;
      PipeCfg cc = c;
      cc.phase_mask = 1 << ph;
      if (ph > 0) {
        // the ticket counters restart for every launch
        NBCTC_CUDA_CHECK(cudaMemsetAsync(c.ctr, 0, 8, stream));
      }
      rc = launch(cc);
      if (rc != NBCTC_OK) return rc;
    }
    return NBCTC_OK;
  }
  return launch(c);
}

}  // namespace nbctc
