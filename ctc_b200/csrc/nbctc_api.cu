// extern "C" surface of libnbctc.so (see include/nbctc.h): argument validation, path
// selection (fused sm_100a kernel vs generic three-kernel path), backward-time rescale,
// and the host-buffer convenience entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"

namespace nbctc {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}
void clear_error() { t_err[0] = 0; }

namespace {

__global__ void scale_grad_kernel(float* __restrict__ g, int64_t n_vec4, int64_t n, int64_t B, int64_t C,
                                  const float* __restrict__ go, int per_seq) {
  if (!per_seq) {
    const float s = go[0];
    if (s == 1.0f) return;  // the common loss.backward() case: nothing to do, no traffic
    float4* g4 = reinterpret_cast<float4*>(g);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec4; i += (int64_t)gridDim.x * blockDim.x) {
      float4 v = g4[i];
      v.x *= s; v.y *= s; v.z *= s; v.w *= s;
      g4[i] = v;
    }
    for (int64_t i = n_vec4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      g[i] *= s;
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      int64_t b = (i / C) % B;
      g[i] *= go[b];
    }
  }
}

// which single-label kernel: NBCTC_FLAG_SEQWARP / NBCTC_FLAG_LOCKSTEP, else the NBCTC_PATH environment variable
// ("seqwarp" | "lockstep", read once), else by batch size (want_seqwarp)
int env_path() {  // 0 none, 2 "lockstep", 3 "seqwarp"
  static const int env = [] {
    const char* v = getenv("NBCTC_PATH");
    return !v ? 0 : v[0] == 'l' ? 2 : v[0] == 's' ? 3 : 0;
  }();
  return env;
}
constexpr uint32_t kPathFlags = NBCTC_FLAG_LOCKSTEP | NBCTC_FLAG_SEQWARP;
// the sequence-per-warp kernel: on request, else for batches that fill the GPU with one sequence per warp (the
// lock-step kernel keeps the small batches: it spreads ONE sequence over a whole CTA)
bool want_seqwarp(uint32_t flags, int64_t T, int64_t B, int64_t C, int64_t Lmax) {
  if (!seqwarp_supported(T, B, C, Lmax)) return false;
  if (flags & kPathFlags) return (flags & NBCTC_FLAG_SEQWARP) != 0;
  if (env_path()) return env_path() == 3;
  // measured cross-overs against the lock-step kernel (gpurun_out/thresholds.log: T=256/512/1024 at the three shape classes)
  return B >= (seqwarp_is_wide(T, B, C, Lmax) ? 384 : Lmax <= 32 ? 2304 : 1536);
}

int check_common(const void* logits, int64_t T, int64_t B, int64_t C, const void* tg, int64_t Lmax,
                 const void* il, const void* tl, const void* loss) {
  if (T < 1 || B < 1 || C < 1 || Lmax < 1) {
    set_error("invalid shape T=%lld B=%lld C=%lld Lmax=%lld", (long long)T, (long long)B, (long long)C, (long long)Lmax);
    return NBCTC_ERR_INVALID_ARG;
  }
  if (!logits || !tg || !il || !tl || !loss) {
    set_error("null pointer argument");
    return NBCTC_ERR_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(logits) & 3) != 0) {
    set_error("logits must be 4-byte aligned");
    return NBCTC_ERR_INVALID_ARG;
  }
  return NBCTC_OK;
}

int run(Problem& p, bool binary, void* ws, size_t ws_bytes, uint32_t flags, cudaStream_t stream, const float* row_lse_in = nullptr,
        float* row_lse_out = nullptr) {
  if (flags & NBCTC_FLAG_NO_GRAD) p.grad = nullptr;
  p.sum_weighted = (flags & NBCTC_FLAG_SUM_WEIGHTED) != 0;
  if (ws != nullptr && (reinterpret_cast<uintptr_t>(ws) & 255) != 0) {
    set_error("workspace must be 256-byte aligned");
    return NBCTC_ERR_INVALID_ARG;
  }
  int rc;
  bool use_sw = !binary && !(flags & NBCTC_FLAG_GENERIC) && want_seqwarp(flags, p.T, p.B, p.C, p.Lmax);
  // the wide-row variant moves rows with TMA: 16-byte aligned tensors only (else the paths below)
  if (use_sw && seqwarp_is_wide(p.T, p.B, p.C, p.Lmax) && !fused_pointers_ok(p)) use_sw = false;
  if ((row_lse_in || row_lse_out) && !use_sw) {
    set_error("row_lse needs the sequence-per-warp kernel (single-label variant, C <= 256, Lmax <= 64)");
    return NBCTC_ERR_UNSUPPORTED;
  }
  if (use_sw) {
    rc = seqwarp_launch(p, ws, ws_bytes, row_lse_in, row_lse_out, stream);
    if (rc != NBCTC_OK) return rc;
    if (p.loss_sum || p.loss_reduced) rc = reduce_loss_launch(p, stream);
    return rc;
  }
  const bool shape_ok = !(flags & NBCTC_FLAG_GENERIC) && fused_supported(p.T, p.B, p.C, p.Lmax, binary);
  if (shape_ok && (flags & NBCTC_FLAG_ALIGNED16) && !fused_pointers_ok(p)) {
    set_error("NBCTC_FLAG_ALIGNED16 was passed but logits / grad_logits are not 16-byte aligned");
    return NBCTC_ERR_INVALID_ARG;
  }
  const bool use_fused = shape_ok && fused_pointers_ok(p);
  if (use_fused) {
    rc = fused_launch(p, binary, ws, ws_bytes, stream);
  } else if (binary && !(flags & NBCTC_FLAG_GENERIC) && tiled_bin_supported(p.T, p.B, p.C, p.Lmax) &&
             (reinterpret_cast<uintptr_t>(p.logits) & 15) == 0) {  // the TMA row copies read 16-byte aligned supersets
    // multi-label: whether the tiled kernels can take the call depends on the target VALUES (exact {0,1}, at most 31
    // classes per state), which only the device sees: workspace = [generic | tiled]; the pre-pass sets a flag and
    // either the tiled kernels or the gated generic kernels do the work
    const size_t g = align_up(generic_workspace_bytes(p.T, p.B, p.C, p.Lmax), 256);
    const size_t t = tiled_bin_workspace_bytes(p.T, p.B, p.C, p.Lmax);
    if (ws == nullptr || ws_bytes < g + t) {
      set_error("workspace too small: need %zu bytes, got %zu", g + t, ws_bytes);
      return NBCTC_ERR_WORKSPACE;
    }
    const int* flag = nullptr;
    rc = tiled_bin_launch(p, static_cast<char*>(ws) + g, t, stream, &flag);
    if (rc == NBCTC_OK) rc = generic_launch(p, true, ws, g, stream, flag);
  } else {
    rc = generic_launch(p, binary, ws, ws_bytes, stream);
  }
  if (rc != NBCTC_OK) return rc;
  if (p.loss_sum || p.loss_reduced) rc = reduce_loss_launch(p, stream);
  return rc;
}

// ---- host-buffer plugin call ------------------------------------------------------------
// The batch is cut into chunks of sequences that travel through three device slots: chunk c+1 is copied in (strided
// 2-D copy out of the (T,B,C) host tensor) while chunk c runs and the gradient of chunk c-1 is copied back, on
// three streams.  Device buffers and streams are kept per device between calls (nbctc_host_release frees them).
struct HostSlot {
  float *x = nullptr, *g = nullptr, *loss = nullptr;
  void* tg = nullptr;
  int64_t *il = nullptr, *tl = nullptr;
  void* ws = nullptr;
  size_t cap_x = 0, cap_g = 0, cap_loss = 0, cap_tg = 0, cap_len = 0, cap_ws = 0;
  cudaEvent_t in_done = nullptr, k_done = nullptr, out_done = nullptr;
};
struct HostCtx {
  bool init = false;
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  HostSlot slot[3];
  std::mutex mu;
};
constexpr int kMaxDevices = 64;
HostCtx g_host[kMaxDevices];

cudaError_t grow(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return cudaSuccess;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  cudaError_t e = cudaMalloc(p, need);
  if (e == cudaSuccess) *cap = need;
  return e;
}

void host_release(HostCtx& h) {
  if (!h.init) return;
  for (HostSlot& s : h.slot) {
    cudaFree(s.x); cudaFree(s.g); cudaFree(s.loss); cudaFree(s.tg); cudaFree(s.il); cudaFree(s.tl); cudaFree(s.ws);
    cudaEventDestroy(s.in_done); cudaEventDestroy(s.k_done); cudaEventDestroy(s.out_done);
    s = HostSlot{};
  }
  cudaStreamDestroy(h.s_in); cudaStreamDestroy(h.s_k); cudaStreamDestroy(h.s_out);
  h.init = false;
}

template <typename TargetT>
int host_entry(int device, bool binary, const float* logits_h, int64_t T, int64_t B, int64_t C,
               const TargetT* tg_h, int64_t Lmax, const int64_t* il_h, const int64_t* tl_h,
               float* loss_h, double* loss_sum_h, float* loss_red_h, float* grad_h, float w_scalar, uint32_t flags) {
  clear_error();
  if (check_common(logits_h, T, B, C, tg_h, Lmax, il_h, tl_h, loss_h) != NBCTC_OK) return NBCTC_ERR_INVALID_ARG;
  if (device < 0 || device >= kMaxDevices) {
    set_error("device index %d out of range", device);
    return NBCTC_ERR_INVALID_ARG;
  }
  NBCTC_CUDA_CHECK(cudaSetDevice(device));
  HostCtx& h = g_host[device];
  std::lock_guard<std::mutex> lock(h.mu);
#define HOST_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                         \
      cudaDeviceSynchronize();                                                           \
      return NBCTC_ERR_CUDA;                                                             \
    }                                                                                    \
  } while (0)
  if (!h.init) {
    HOST_TRY(cudaStreamCreateWithFlags(&h.s_in, cudaStreamNonBlocking));
    HOST_TRY(cudaStreamCreateWithFlags(&h.s_k, cudaStreamNonBlocking));
    HOST_TRY(cudaStreamCreateWithFlags(&h.s_out, cudaStreamNonBlocking));
    for (HostSlot& s : h.slot) {
      HOST_TRY(cudaEventCreateWithFlags(&s.in_done, cudaEventDisableTiming));
      HOST_TRY(cudaEventCreateWithFlags(&s.k_done, cudaEventDisableTiming));
      HOST_TRY(cudaEventCreateWithFlags(&s.out_done, cudaEventDisableTiming));
    }
    h.init = true;
  }
  const bool want_grad = grad_h != nullptr && !(flags & NBCTC_FLAG_NO_GRAD);
  // chunk of sequences: ~64 MB of logits, a multiple of 4 sequences (16-byte aligned slabs for any C), >= 16 chunks
  // for a large batch so that the copies of neighbouring chunks hide the kernel
  const int64_t row_bytes = T * C * 4;
  int64_t Bc = std::max<int64_t>(4, (int64_t)(64e6 / (double)row_bytes) / 4 * 4);
  static const int min_chunks = [] { const char* v = getenv("NBCTC_HOST_CHUNKS"); return v ? std::max(1, atoi(v)) : 16; }();
  Bc = std::min(Bc, std::max<int64_t>(4, ((B + min_chunks - 1) / min_chunks + 3) / 4 * 4));
  Bc = std::min(Bc, B);
  const int64_t nchunk = (B + Bc - 1) / Bc;
  const size_t tg_per_seq = (size_t)Lmax * (binary ? (size_t)C : 1) * sizeof(TargetT);
  const size_t ws_bytes = nbctc_workspace_bytes(T, Bc, C, Lmax, binary ? 1 : 0, flags);
  const int nslot = (int)std::min<int64_t>(3, nchunk);
  for (int i = 0; i < nslot; ++i) {
    HostSlot& s = h.slot[i];
    HOST_TRY(grow((void**)&s.x, &s.cap_x, (size_t)T * Bc * C * 4));
    if (want_grad) HOST_TRY(grow((void**)&s.g, &s.cap_g, (size_t)T * Bc * C * 4));
    HOST_TRY(grow((void**)&s.loss, &s.cap_loss, (size_t)Bc * 4));
    HOST_TRY(grow(&s.tg, &s.cap_tg, tg_per_seq * Bc));
    size_t cap_len2 = s.cap_len;
    HOST_TRY(grow((void**)&s.il, &s.cap_len, (size_t)Bc * 8));
    HOST_TRY(grow((void**)&s.tl, &cap_len2, (size_t)Bc * 8));
    HOST_TRY(grow(&s.ws, &s.cap_ws, ws_bytes ? ws_bytes : 256));
  }
  int rc = NBCTC_OK;
  for (int64_t c = 0; c < nchunk && rc == NBCTC_OK; ++c) {
    HostSlot& s = h.slot[c % 3];
    const int64_t b0 = c * Bc, nb = std::min(Bc, B - b0);
    // the slot's previous occupant (chunk c-3) must have left the device
    HOST_TRY(cudaStreamWaitEvent(h.s_in, s.out_done, 0));
    HOST_TRY(cudaMemcpy2DAsync(s.x, (size_t)nb * C * 4, logits_h + b0 * C, (size_t)B * C * 4, (size_t)nb * C * 4, (size_t)T,
                               cudaMemcpyHostToDevice, h.s_in));
    HOST_TRY(cudaMemcpyAsync(s.tg, reinterpret_cast<const char*>(tg_h) + tg_per_seq * b0, tg_per_seq * nb, cudaMemcpyHostToDevice, h.s_in));
    HOST_TRY(cudaMemcpyAsync(s.il, il_h + b0, (size_t)nb * 8, cudaMemcpyHostToDevice, h.s_in));
    HOST_TRY(cudaMemcpyAsync(s.tl, tl_h + b0, (size_t)nb * 8, cudaMemcpyHostToDevice, h.s_in));
    HOST_TRY(cudaEventRecord(s.in_done, h.s_in));
    HOST_TRY(cudaStreamWaitEvent(h.s_k, s.in_done, 0));
    Problem p{};
    p.logits = s.x;
    if (binary) p.targets = reinterpret_cast<const float*>(s.tg); else p.labels = reinterpret_cast<const int32_t*>(s.tg);
    p.in_len = s.il; p.tgt_len = s.tl; p.loss = s.loss; p.loss_sum = nullptr; p.loss_reduced = nullptr;
    p.grad = want_grad ? s.g : nullptr; p.seq_w = nullptr; p.w_scalar = w_scalar;
    p.T = T; p.B = nb; p.C = C; p.Lmax = Lmax;
    rc = run(p, binary, s.ws, s.cap_ws, flags, h.s_k);
    if (rc != NBCTC_OK) break;
    HOST_TRY(cudaEventRecord(s.k_done, h.s_k));
    HOST_TRY(cudaStreamWaitEvent(h.s_out, s.k_done, 0));
    HOST_TRY(cudaMemcpyAsync(loss_h + b0, s.loss, (size_t)nb * 4, cudaMemcpyDeviceToHost, h.s_out));
    if (want_grad)
      HOST_TRY(cudaMemcpy2DAsync(grad_h + b0 * C, (size_t)B * C * 4, s.g, (size_t)nb * C * 4, (size_t)nb * C * 4, (size_t)T,
                                 cudaMemcpyDeviceToHost, h.s_out));
    HOST_TRY(cudaEventRecord(s.out_done, h.s_out));
  }
  HOST_TRY(cudaStreamSynchronize(h.s_in));
  HOST_TRY(cudaStreamSynchronize(h.s_k));
  HOST_TRY(cudaStreamSynchronize(h.s_out));
#undef HOST_TRY
  if (rc != NBCTC_OK) return rc;
  // the reduction of NoBlankCTC.py:140 over the whole batch, float64, ascending b
  if (loss_sum_h || loss_red_h) {
    double acc = 0.0;
    for (int64_t b = 0; b < B; ++b) acc += (double)loss_h[b];
    if (loss_sum_h) *loss_sum_h = acc;
    if (loss_red_h) *loss_red_h = (float)((double)w_scalar * acc);
  }
  return rc;
}

}  // namespace
}  // namespace nbctc

using namespace nbctc;

extern "C" {

int nbctc_version(void) { return NBCTC_VERSION; }

const char* nbctc_last_error(void) { return t_err; }

uint64_t nbctc_kernel_launch_count(void) { return g_launch_count.load(std::memory_order_relaxed); }

size_t nbctc_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax, int binary, uint32_t flags) {
  if (T < 1 || B < 1 || C < 1 || Lmax < 1) return 0;
  size_t g = generic_workspace_bytes(T, B, C, Lmax);
  if (flags & NBCTC_FLAG_GENERIC) return g;
  // the generic size is kept as a floor: a call with 16-byte misaligned tensors falls back to that path
  size_t f = 0;
  if (fused_supported(T, B, C, Lmax, binary != 0)) f = fused_workspace_bytes(T, B, C, Lmax, binary != 0);
  if (!binary && want_seqwarp(flags, T, B, C, Lmax)) {
    const size_t sw = seqwarp_workspace_bytes(T, B, C, Lmax);
    if (!seqwarp_is_wide(T, B, C, Lmax)) return sw;  // any alignment
    return std::max(sw, f ? ((flags & NBCTC_FLAG_ALIGNED16) ? f : std::max(g, f)) : g);
  }
  if (f) return (flags & NBCTC_FLAG_ALIGNED16) ? f : std::max(g, f);
  if (binary && tiled_bin_supported(T, B, C, Lmax)) return align_up(g, 256) + tiled_bin_workspace_bytes(T, B, C, Lmax);
  return g;
}

int nbctc_loss_grad_f32(const float* logits, int64_t T, int64_t B, int64_t C, const int32_t* labels, int64_t Lmax,
                        const int64_t* input_lengths, const int64_t* target_lengths, float* loss_per_seq,
                        double* loss_sum, float* loss_reduced, float* grad_logits, const float* seq_weights, float weight_scalar,
                        void* workspace, size_t workspace_bytes, uint32_t flags, nbctc_stream_t stream) {
  clear_error();
  int rc = check_common(logits, T, B, C, labels, Lmax, input_lengths, target_lengths, loss_per_seq);
  if (rc != NBCTC_OK) return rc;
  Problem p{};
  p.logits = logits; p.labels = labels; p.targets = nullptr;
  p.in_len = input_lengths; p.tgt_len = target_lengths;
  p.loss = loss_per_seq; p.loss_sum = loss_sum; p.loss_reduced = loss_reduced; p.grad = grad_logits;
  p.seq_w = seq_weights; p.w_scalar = weight_scalar;
  p.T = T; p.B = B; p.C = C; p.Lmax = Lmax;
  return run(p, false, workspace, workspace_bytes, flags, static_cast<cudaStream_t>(stream));
}

int nbctc_loss_grad_lse_f32(const float* logits, int64_t T, int64_t B, int64_t C, const int32_t* labels, int64_t Lmax,
                            const int64_t* input_lengths, const int64_t* target_lengths, const float* row_lse_in, float* row_lse_out,
                            float* loss_per_seq, double* loss_sum, float* loss_reduced, float* grad_logits, const float* seq_weights,
                            float weight_scalar, void* workspace, size_t workspace_bytes, uint32_t flags, nbctc_stream_t stream) {
  clear_error();
  int rc = check_common(logits, T, B, C, labels, Lmax, input_lengths, target_lengths, loss_per_seq);
  if (rc != NBCTC_OK) return rc;
  Problem p{};
  p.logits = logits; p.labels = labels; p.targets = nullptr;
  p.in_len = input_lengths; p.tgt_len = target_lengths;
  p.loss = loss_per_seq; p.loss_sum = loss_sum; p.loss_reduced = loss_reduced; p.grad = grad_logits;
  p.seq_w = seq_weights; p.w_scalar = weight_scalar;
  p.T = T; p.B = B; p.C = C; p.Lmax = Lmax;
  return run(p, false, workspace, workspace_bytes, (flags & ~kPathFlags) | NBCTC_FLAG_SEQWARP, static_cast<cudaStream_t>(stream),
             row_lse_in, row_lse_out);
}

int nbbctc_loss_grad_f32(const float* logits, int64_t T, int64_t B, int64_t C, const float* targets, int64_t Lmax,
                         const int64_t* input_lengths, const int64_t* target_lengths, float* loss_per_seq,
                         double* loss_sum, float* loss_reduced, float* grad_logits, const float* seq_weights, float weight_scalar,
                         void* workspace, size_t workspace_bytes, uint32_t flags, nbctc_stream_t stream) {
  clear_error();
  int rc = check_common(logits, T, B, C, targets, Lmax, input_lengths, target_lengths, loss_per_seq);
  if (rc != NBCTC_OK) return rc;
  Problem p{};
  p.logits = logits; p.labels = nullptr; p.targets = targets;
  p.in_len = input_lengths; p.tgt_len = target_lengths;
  p.loss = loss_per_seq; p.loss_sum = loss_sum; p.loss_reduced = loss_reduced; p.grad = grad_logits;
  p.seq_w = seq_weights; p.w_scalar = weight_scalar;
  p.T = T; p.B = B; p.C = C; p.Lmax = Lmax;
  return run(p, true, workspace, workspace_bytes, flags, static_cast<cudaStream_t>(stream));
}

int nbctc_aux_ce_f32(const float* logits, int64_t T, int64_t B, int64_t C, const int64_t* frame_index,
                     const int64_t* input_lengths, const int32_t* class_index, const float* multi_hot, float alpha_weight,
                     const float* seq_weights, float* ce_per_seq, float* grad_logits, nbctc_stream_t stream) {
  clear_error();
  if (T < 1 || B < 1 || C < 1 || !logits || !ce_per_seq || (!frame_index && !input_lengths)) {
    set_error("invalid argument to nbctc_aux_ce_f32");
    return NBCTC_ERR_INVALID_ARG;
  }
  if ((class_index == nullptr) == (multi_hot == nullptr)) {
    set_error("nbctc_aux_ce_f32 takes exactly one of class_index and multi_hot");
    return NBCTC_ERR_INVALID_ARG;
  }
  return aux_ce_launch(logits, T, B, C, frame_index, input_lengths, class_index, multi_hot, class_index ? 0 : 1, alpha_weight,
                       seq_weights, ce_per_seq, grad_logits, static_cast<cudaStream_t>(stream));
}

int nbctc_scale_grad_f32(float* grad_logits, int64_t T, int64_t B, int64_t C, const float* grad_out, int per_seq,
                         nbctc_stream_t stream) {
  clear_error();
  if (!grad_logits || !grad_out || T < 1 || B < 1 || C < 1) {
    set_error("invalid argument to nbctc_scale_grad_f32");
    return NBCTC_ERR_INVALID_ARG;
  }
  const int64_t n = T * B * C;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(grad_logits) & 15) == 0;
  const int64_t n4 = vec_ok ? n / 4 : 0;
  const int threads = 256;
  const int64_t want = (n / 4 + threads - 1) / threads;
  const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, 148 * 16));
  scale_grad_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(grad_logits, n4, n, B, C, grad_out, per_seq);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

int nbctc_host_release(int device) {
  clear_error();
  if (device < 0 || device >= kMaxDevices) return NBCTC_ERR_INVALID_ARG;
  HostCtx& h = g_host[device];
  std::lock_guard<std::mutex> lock(h.mu);
  if (h.init) {
    NBCTC_CUDA_CHECK(cudaSetDevice(device));
    host_release(h);
  }
  return NBCTC_OK;
}

int nbctc_loss_grad_host_f32(int device, const float* logits_host, int64_t T, int64_t B, int64_t C,
                             const int32_t* labels_host, int64_t Lmax, const int64_t* input_lengths_host,
                             const int64_t* target_lengths_host, float* loss_per_seq_host, double* loss_sum_host,
                             float* loss_reduced_host, float* grad_logits_host, float weight_scalar, uint32_t flags) {
  return host_entry<int32_t>(device, false, logits_host, T, B, C, labels_host, Lmax, input_lengths_host,
                             target_lengths_host, loss_per_seq_host, loss_sum_host, loss_reduced_host, grad_logits_host, weight_scalar, flags);
}

int nbbctc_loss_grad_host_f32(int device, const float* logits_host, int64_t T, int64_t B, int64_t C,
                              const float* targets_host, int64_t Lmax, const int64_t* input_lengths_host,
                              const int64_t* target_lengths_host, float* loss_per_seq_host, double* loss_sum_host,
                              float* loss_reduced_host, float* grad_logits_host, float weight_scalar, uint32_t flags) {
  return host_entry<float>(device, true, logits_host, T, B, C, targets_host, Lmax, input_lengths_host,
                           target_lengths_host, loss_per_seq_host, loss_sum_host, loss_reduced_host, grad_logits_host, weight_scalar, flags);
}

}  // extern "C"
