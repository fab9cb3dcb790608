// Fused no-blank CTC forward+backward for sm_100a: ONE kernel reads the logits once from HBM, writes the
// gradient once, and keeps everything in between on chip (kernel in fused_kernel.cuh).
//
//   * row warps stream the (t,b) rows HBM -> registers with 128-bit loads on the 16-byte-aligned superset of
//     each (only 4-byte aligned, C=157) row.  Phase 1: row log-partition (NoBlankCTC.py:136) + per-state
//     emission gather (NoBlankCTC.py:96-102) -> p-tiles in shared memory; loads carry an L2 evict_last policy.
//     Phase 2: the same rows again, in reverse time order so the most recently read rows are re-read first
//     (L2 hits, evict_first) -> w*(softmax - scatter(gamma)) -> 128-bit evict_first stores.
//   * one chain warp per sequence runs the lattice recursions (NoBlankCTC.py:71-87) in the LINEAR domain in
//     float64 with exact power-of-two rescaling once per tile: a step is a shuffle, an add and a multiply, and
//     sum_s alpha_t(s) beta_t(s) = Z holds to 1e-13 so gamma needs no per-row normalisation.  Phase 1 stores
//     one alpha checkpoint per tile of kTT steps; phase 2 replays alpha inside the tile next to the beta
//     recursion (the reference's backward pass is commented out at NoBlankCTC.py:113-125; autograd does it).
//   * tiles are handed over through shared memory with mbarriers (full/empty per ring buffer).
//
// This file: shape -> launch plan (states per lane, lanes per row, shared-memory carve-up), workspace, dispatch.
// Algorithmic HBM bytes per sequence: 2*4*T*C (+ labels); per real lattice cell 8*C/mean(L).
#include <algorithm>

#include "fused_kernel.cuh"

namespace nbctc {
#ifdef NBCTC_PROF
namespace fused { __device__ long long* g_nbctc_prof = nullptr; }
extern "C" int nbctc_debug_set_prof(long long* dev_buf) {
  return (int)cudaMemcpyToSymbol(fused::g_nbctc_prof, &dev_buf, sizeof(dev_buf));
}
#endif
namespace {

struct Plan {
  bool ok;
  FusedCfg cfg;
};

Plan make_plan(int64_t T, int64_t B, int64_t C, int64_t Lmax, const void* logits_ptr) {
  (void)B;
  Plan pl{};
  pl.ok = false;
  if (Lmax > 256 || T > ((int64_t)1 << 28) || C > ((int64_t)1 << 22)) return pl;
  FusedCfg& c = pl.cfg;
  c.NS = Lmax <= 32 ? 2 : Lmax <= 64 ? 4 : Lmax <= 128 ? 8 : 16;  // 16 chain lanes per direction
  c.Lpad = 16 * c.NS;
  // chunks per row: rows of a tensor with C % 4 == 0 all share the base pointer's alignment
  const int off_base = (int)((reinterpret_cast<uintptr_t>(logits_ptr) >> 2) & 3);
  const int64_t nch = (C % 4 == 0) ? (off_base + C + 3) / 4 : (3 + C + 3) / 4;
  c.NSEG = 1;
  if (nch <= 16) {
    c.LPR = 4;
    c.CPL = (int)((nch + 3) / 4);
  } else if (nch <= 64) {
    c.LPR = 8;
    c.CPL = (int)std::max<int64_t>(3, (nch + 7) / 8);
  } else {
    // long rows: one row per warp pass (a single scatter buffer per warp), 32 lanes x CPL chunks per segment
    c.LPR = 32;
    c.CPL = nch <= 96 ? 3 : nch <= 128 ? 4 : nch <= 192 ? 6 : 8;
    c.NSEG = (int)((nch + 32 * c.CPL - 1) / (32 * c.CPL));
  }
  c.RS = (int)(16 * nch);
  c.logits_end = nullptr;
  c.NTmax = (int)((T + kTT - 1) / kTT);
  const size_t budget = 200 * 1024;
  auto layout = [&](int nslot, bool ck_glob, bool lse_glob) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 16); return (uint32_t)o; };
    c.NSLOT = nslot; c.NBUFP = c.Lpad <= 64 ? 2 * kNW : kNW + 1; c.NBUFG = kNW; c.ckpt_global = ck_glob;
    c.lse_global = lse_glob;
    const int nbuf = c.NBUFP;
    c.o_bar = take(sizeof(uint64_t) * (3 * kMaxBuf + 2 * kMaxSlot) + 16);
    c.o_lab = take(sizeof(int) * c.Lpad);
    c.o_lse = take(lse_glob ? 16 : sizeof(float) * T);
    c.o_ckpt = take(ck_glob ? 16 : sizeof(double) * (size_t)c.NTmax * c.Lpad);
    c.o_cke = take(ck_glob ? 16 : sizeof(int) * (size_t)c.NTmax);
    c.o_ptile = take(sizeof(float) * (size_t)nbuf * kTT * c.Lpad);
    c.o_gtile = take(sizeof(float) * (size_t)c.NBUFG * kTT * c.Lpad);
    c.o_abtile = take(sizeof(double) * 2 * (size_t)kTT * c.Lpad);
    off = align_up(off, 128);
    c.o_ring = take((size_t)nslot * kTT * c.RS);
    c.smem_bytes = (uint32_t)off;
    return off;
  };
  // pass 0: everything in shared memory and <= 55 KB so that 4 CTAs share an SM; pass 1: <= 74 KB (3 CTAs);
  // later passes move the checkpoints / row constants to the workspace and let one CTA take up to ~200 KB.
  bool placed = false;
  for (int pass = 0; pass < 4 && !placed; ++pass) {
    const bool ckg = pass >= 2, lsg = pass >= 3;
    const size_t cap = pass == 0 ? 55 * 1024 : pass == 1 ? 74 * 1024 : budget;
    const int lo = pass <= 1 ? 5 : 3;
    for (int nslot = 7; nslot >= lo && !placed; --nslot)
      if (layout(nslot, ckg, lsg) <= cap) placed = true;
  }
  if (!placed) return pl;
  pl.ok = true;
  return pl;
}

size_t plan_ws_bytes(const Plan& pl, int64_t T, int64_t B) {
  size_t off = 256;
  if (!pl.ok) return off;
  if (pl.cfg.ckpt_global) {
    off = align_up(off + sizeof(double) * (size_t)B * pl.cfg.NTmax * pl.cfg.Lpad, 256);
    off = align_up(off + sizeof(int) * (size_t)B * pl.cfg.NTmax, 256);
  }
  if (pl.cfg.lse_global) off = align_up(off + sizeof(float) * (size_t)B * T, 256);
  return off;
}

}  // namespace

bool fused_supported(int64_t T, int64_t B, int64_t C, int64_t Lmax, bool binary) {
  if (binary) return false;
  // must not depend on the (unknown here) base alignment: plan for both extremes
  return make_plan(T, B, C, Lmax, reinterpret_cast<const void*>(uintptr_t(0))).ok &&
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
  pl.cfg.logits_end = p.logits + p.T * p.B * p.C;
  switch (pl.cfg.NS) {
    case 2: return launch_fused_ns2(p, pl.cfg, stream);
    case 4: return launch_fused_ns4(p, pl.cfg, stream);
    case 8: return launch_fused_ns8(p, pl.cfg, stream);
    default: return launch_fused_ns16(p, pl.cfg, stream);
  }
}

}  // namespace nbctc
