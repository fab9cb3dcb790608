// Fused no-blank CTC forward+backward for sm_100a: ONE kernel reads the logits from HBM once, writes the
// gradient once, and keeps everything in between on chip or in L2 (kernel in stream_kernel.cuh).
//
//   * a CTA owns GB batch-adjacent sequences; their rows at one time step are contiguous in (T,B,C) ("slab").
//     Row warp w owns time step w of every tile of TT steps: one TMA bulk copy brings its slab into a
//     shared-memory ring, the warp computes the row log-partitions (NoBlankCTC.py:136), turns the slab into
//     w*softmax in place, gathers the per-state emissions (NoBlankCTC.py:96-102) and sends the slab to the
//     gradient tensor with one TMA bulk store.
//   * one chain warp per sequence runs the lattice recursions (NoBlankCTC.py:71-87) in the LINEAR domain in
//     float64 with exact power-of-two rescaling once per tile: a step is a shuffle, an add and a multiply, and
//     sum_s alpha_t(s) beta_t(s) = Z holds to 1e-13 so gamma needs no per-row normalisation.  Phase 1 stores
//     one alpha checkpoint per tile; phase 2 walks the tiles downwards, replays alpha inside the tile next to
//     the beta recursion (the reference's backward pass is commented out at NoBlankCTC.py:113-125; autograd
//     does it), and the row warps add -w*gamma to the gradient rows in shared memory, in conflict-free rank rounds.
//   * the two roles advance in lock step, one __syncthreads() per tile; nobody polls.
//
// This file: shape -> launch plan (states per lane, lanes per row, group size, shared-memory carve-up),
// workspace, dispatch.  Algorithmic HBM bytes per sequence: 2*4*T*C (+ labels); per real lattice cell 8*C/mean(L).
#include <stdlib.h>

#include <algorithm>

#include "stream_kernel.cuh"

namespace nbctc {
long long* g_stream_prof = nullptr;
namespace {

struct Plan {
  bool ok;
  StreamCfg cfg;
};

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

Plan make_plan(int64_t T, int64_t B, int64_t C, int64_t Lmax) {
  (void)B;
  Plan pl{};
  pl.ok = false;
  // C >= 4: a 16-byte chunk of a slab then touches at most two rows (grad_step)
  if (Lmax > 256 || T > ((int64_t)1 << 28) || C > ((int64_t)1 << 22) || C < 4) return pl;
  StreamCfg& c = pl.cfg;
  c.NS = Lmax <= 32 ? 2 : Lmax <= 64 ? 4 : Lmax <= 128 ? 8 : 16;  // 16 chain lanes per direction
  c.Lpad = 16 * c.NS;
  c.TT = c.NS >= 16 ? 2 : 8;
  const int PS = c.Lpad + 8, AS = c.Lpad + 8;
  const int PSEQ = c.TT * PS + 8, ABSEQ = 2 * c.TT * AS + 8;
  // chunks per row: with C % 4 == 0 every row starts on a 16-byte boundary, otherwise at any of the 4 phases
  const int64_t nch = (C % 4 == 0) ? C / 4 : (3 + C + 3) / 4;
  // lanes per row (LPR) fixes the sequences per CTA (GB = 32 / LPR).  NBCTC_LPR / NBCTC_CTAS override (tuning).
  const int want_lpr = env_int("NBCTC_LPR", 0);
  c.ctas_per_sm = std::max(1, env_int("NBCTC_CTAS", 1));
  c.NSEG = 1;
  c.LPR = want_lpr ? want_lpr : nch <= 16 ? 4 : nch <= 64 ? 8 : 32;
  if (c.LPR == 4) {
    c.CPL = (int)((nch + 3) / 4);
    if (c.CPL > 4) return pl;
  } else if (c.LPR == 8) {
    c.CPL = (int)std::max<int64_t>(3, (nch + 7) / 8);
    if (c.CPL > 8) return pl;
  } else if (c.LPR == 16) {
    c.CPL = (int)std::max<int64_t>(2, (nch + 15) / 16);
    if (c.CPL > 4) return pl;
  } else if (c.LPR == 32) {
    c.CPL = nch <= 96 ? 3 : nch <= 128 ? 4 : nch <= 192 ? 6 : 8;
    c.NSEG = (int)((nch + 32 * c.CPL - 1) / (32 * c.CPL));
  } else {
    return pl;
  }
  c.GB = 32 / c.LPR;
  c.NTmax = (int)((T + c.TT - 1) / c.TT);
  const size_t cap = (size_t)(c.ctas_per_sm >= 2 ? 113 : 227) * 1024;
  auto layout = [&](bool ck_glob) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 16); return (uint32_t)o; };
    const int gb = c.GB;
    c.ckpt_global = ck_glob;
    c.RSg = (int)align_up((size_t)gb * C * 4, 16) + 32;
    c.o_bar = take(sizeof(uint64_t) * kNSlot * c.TT);
    c.o_info = take(sizeof(int) * 4 * kMaxGB);
    c.o_lab = take(sizeof(int) * gb * c.Lpad);
    c.o_ckpt = take(ck_glob ? 16 : sizeof(double) * (size_t)gb * c.NTmax * c.Lpad);
    c.o_cke = take(ck_glob ? 16 : sizeof(int) * (size_t)gb * c.NTmax * (c.NS > 2 ? 32 : 16));
    c.o_ptile = take(sizeof(float) * 2 * (size_t)gb * PSEQ);
    c.o_s2 = take(sizeof(double) * 2 * gb);
    c.o_pub = take(16 * (size_t)gb);
    c.o_ab = take(sizeof(double) * 2 * (size_t)gb * ABSEQ);
    off = align_up(off, 128);
    c.o_ring = take((size_t)kNSlot * c.TT * c.RSg);
    c.smem_bytes = (uint32_t)off;
    return off;
  };
  // checkpoints on chip if they fit next to the ring, else in the workspace
  bool placed = false;
  for (int pass = 0; pass < 2 && !placed; ++pass)
    if (layout(pass >= 1) <= cap) placed = true;
  if (!placed) return pl;
  pl.ok = true;
  return pl;
}

// workspace: [256 | alpha checkpoints | their lane scales | repair: flags [B], row log-partitions [B][T], alpha
// checkpoints of the log-domain repair kernel [B][ceil(T/4)][Lpad]]
struct WsLayout {
  size_t o_ckpt, o_cke, o_flag, o_rlse, o_rckx, total;
  int64_t rK;
};
WsLayout ws_layout(const Plan& pl, int64_t T, int64_t B) {
  WsLayout l{};
  size_t off = 256;
  l.o_ckpt = off;
  off = align_up(off + sizeof(double) * (size_t)B * pl.cfg.NTmax * pl.cfg.Lpad, 256);
  l.o_cke = off;
  off = align_up(off + sizeof(int) * (size_t)B * pl.cfg.NTmax * (pl.cfg.NS > 2 ? 32 : 16), 256);
  l.o_flag = off;
  off = align_up(off + sizeof(int) * (size_t)B, 256);
  l.o_rlse = off;
  off = align_up(off + sizeof(float) * (size_t)B * T, 256);
  l.rK = (T + 3) / 4;
  l.o_rckx = off;
  off = align_up(off + sizeof(double) * (size_t)B * l.rK * pl.cfg.Lpad, 256);
  l.total = off;
  return l;
}
size_t plan_ws_bytes(const Plan& pl, int64_t T, int64_t B) {
  if (!pl.ok) return 256;
  return ws_layout(pl, T, B).total;
}

}  // namespace

bool fused_supported(int64_t T, int64_t B, int64_t C, int64_t Lmax, bool binary) {
  if (binary) return false;
  return make_plan(T, B, C, Lmax).ok;
}

size_t fused_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax, bool binary) {
  (void)binary;
  Plan a = make_plan(T, B, C, Lmax);
  if (!a.ok) return 256;
  return plan_ws_bytes(a, T, B);  // sized for checkpoints in the workspace whatever the plan (NBCTC_CTAS may differ)
}

bool fused_pointers_ok(const Problem& p) {
  // the bulk copies move 16-byte aligned supersets of the rows and the ring keeps the global 16-byte phase
  return (reinterpret_cast<uintptr_t>(p.logits) & 15) == 0 && (p.grad == nullptr || (reinterpret_cast<uintptr_t>(p.grad) & 15) == 0);
}

int fused_launch(const Problem& p, bool binary, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (binary) {
    set_error("fused binary path not available");
    return NBCTC_ERR_UNSUPPORTED;
  }
  Plan pl = make_plan(p.T, p.B, p.C, p.Lmax);
  if (!pl.ok || !fused_pointers_ok(p)) {
    set_error("shape or pointer alignment not supported by the fused kernel");
    return NBCTC_ERR_UNSUPPORTED;
  }
  const WsLayout l = ws_layout(pl, p.T, p.B);
  if (ws == nullptr || ws_bytes < l.total) {
    set_error("workspace too small: need %zu bytes, got %zu", l.total, ws_bytes);
    return NBCTC_ERR_WORKSPACE;
  }
  char* w = static_cast<char*>(ws);
  if (pl.cfg.ckpt_global) {
    pl.cfg.ws_ckpt = reinterpret_cast<double*>(w + l.o_ckpt);
    pl.cfg.ws_cke = reinterpret_cast<int*>(w + l.o_cke);
  }
  pl.cfg.floor_flag = reinterpret_cast<int*>(w + l.o_flag);
  pl.cfg.prof = g_stream_prof;
  int rc;
  switch (pl.cfg.NS) {
    case 2: rc = launch_stream_ns2(p, pl.cfg, stream); break;
    case 4: rc = launch_stream_ns4(p, pl.cfg, stream); break;
    case 8: rc = launch_stream_ns8(p, pl.cfg, stream); break;
    default: rc = launch_stream_ns16(p, pl.cfg, stream); break;
  }
  if (rc != NBCTC_OK) return rc;
  // sequences with an emission below the float32 floor are redone in the log domain (nbctc_logdom.cu)
  LogWs lw{pl.cfg.floor_flag, reinterpret_cast<float*>(w + l.o_rlse), p.T, reinterpret_cast<double*>(w + l.o_rckx), l.rK * pl.cfg.Lpad};
  return logdom_repair_launch(p, lw, stream);
}

}  // namespace nbctc

// role-profiler hook (only meaningful in -DNBCTC_PROF builds; harmless otherwise)
extern "C" int nbctc_debug_set_prof(long long* dev_buf) {
  nbctc::g_stream_prof = dev_buf;
  return 0;
}
