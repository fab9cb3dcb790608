// extern "C" surface of libnbctc.so (see include/nbctc.h): argument validation, path
// selection (fused sm_100a kernel vs generic three-kernel path), backward-time rescale,
// and the host-buffer convenience entry points.
#include <stdarg.h>
#include <string.h>

#include <algorithm>

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

int run(Problem& p, bool binary, void* ws, size_t ws_bytes, uint32_t flags, cudaStream_t stream) {
  if (flags & NBCTC_FLAG_NO_GRAD) p.grad = nullptr;
  int rc;
  const bool shape_ok = !(flags & NBCTC_FLAG_GENERIC) && fused_supported(p.T, p.B, p.C, p.Lmax, binary);
  if (shape_ok && (flags & NBCTC_FLAG_ALIGNED16) && !fused_pointers_ok(p)) {
    set_error("NBCTC_FLAG_ALIGNED16 was passed but logits / grad_logits are not 16-byte aligned");
    return NBCTC_ERR_INVALID_ARG;
  }
  const bool use_fused = shape_ok && fused_pointers_ok(p);
  const bool use_pipe = !binary && !(flags & (NBCTC_FLAG_GENERIC | NBCTC_FLAG_LOCKSTEP)) && fused_pointers_ok(p) &&
                        pipe_supported(p.T, p.B, p.C, p.Lmax);
  if (use_pipe) {
    rc = pipe_launch(p, ws, ws_bytes, stream);
  } else if (use_fused) {
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

template <typename TargetT>
int host_entry(int device, bool binary, const float* logits_h, int64_t T, int64_t B, int64_t C,
               const TargetT* tg_h, int64_t Lmax, const int64_t* il_h, const int64_t* tl_h,
               float* loss_h, double* loss_sum_h, float* loss_red_h, float* grad_h, float w_scalar, uint32_t flags) {
  clear_error();
  if (check_common(logits_h, T, B, C, tg_h, Lmax, il_h, tl_h, loss_h) != NBCTC_OK) return NBCTC_ERR_INVALID_ARG;
  NBCTC_CUDA_CHECK(cudaSetDevice(device));
  cudaStream_t st;
  NBCTC_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const size_t n = (size_t)T * B * C;
  const size_t tg_elems = binary ? (size_t)B * Lmax * C : (size_t)B * Lmax;
  const size_t ws_bytes = nbctc_workspace_bytes(T, B, C, Lmax, binary ? 1 : 0, flags);
  float *d_x = nullptr, *d_g = nullptr, *d_loss = nullptr;
  TargetT* d_tg = nullptr;
  int64_t *d_il = nullptr, *d_tl = nullptr;
  double* d_sum = nullptr;
  float* d_red = nullptr;
  void* d_ws = nullptr;
  int rc = NBCTC_OK;
  auto cleanup = [&]() {
    cudaFree(d_x); cudaFree(d_g); cudaFree(d_loss); cudaFree(d_tg); cudaFree(d_il); cudaFree(d_tl);
    cudaFree(d_sum); cudaFree(d_red); cudaFree(d_ws);
    cudaStreamDestroy(st);
  };
#define HOST_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                         \
      cleanup();                                                                         \
      return NBCTC_ERR_CUDA;                                                             \
    }                                                                                    \
  } while (0)
  const bool want_grad = grad_h != nullptr && !(flags & NBCTC_FLAG_NO_GRAD);
  HOST_TRY(cudaMalloc(&d_x, n * sizeof(float)));
  if (want_grad) HOST_TRY(cudaMalloc(&d_g, n * sizeof(float)));
  HOST_TRY(cudaMalloc(&d_loss, B * sizeof(float)));
  HOST_TRY(cudaMalloc(&d_tg, tg_elems * sizeof(TargetT)));
  HOST_TRY(cudaMalloc(&d_il, B * sizeof(int64_t)));
  HOST_TRY(cudaMalloc(&d_tl, B * sizeof(int64_t)));
  HOST_TRY(cudaMalloc(&d_sum, sizeof(double)));
  HOST_TRY(cudaMalloc(&d_red, sizeof(float)));
  HOST_TRY(cudaMalloc(&d_ws, ws_bytes ? ws_bytes : 256));
  HOST_TRY(cudaMemcpyAsync(d_x, logits_h, n * sizeof(float), cudaMemcpyHostToDevice, st));
  HOST_TRY(cudaMemcpyAsync(d_tg, tg_h, tg_elems * sizeof(TargetT), cudaMemcpyHostToDevice, st));
  HOST_TRY(cudaMemcpyAsync(d_il, il_h, B * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  HOST_TRY(cudaMemcpyAsync(d_tl, tl_h, B * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  Problem p{};
  p.logits = d_x;
  if (binary) p.targets = reinterpret_cast<const float*>(d_tg); else p.labels = reinterpret_cast<const int32_t*>(d_tg);
  p.in_len = d_il; p.tgt_len = d_tl; p.loss = d_loss; p.loss_sum = d_sum; p.loss_reduced = d_red;
  p.grad = want_grad ? d_g : nullptr; p.seq_w = nullptr; p.w_scalar = w_scalar;
  p.T = T; p.B = B; p.C = C; p.Lmax = Lmax;
  rc = run(p, binary, d_ws, ws_bytes, flags, st);
  if (rc == NBCTC_OK) {
    HOST_TRY(cudaMemcpyAsync(loss_h, d_loss, B * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (loss_sum_h) HOST_TRY(cudaMemcpyAsync(loss_sum_h, d_sum, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (loss_red_h) HOST_TRY(cudaMemcpyAsync(loss_red_h, d_red, sizeof(float), cudaMemcpyDeviceToHost, st));
    if (want_grad) HOST_TRY(cudaMemcpyAsync(grad_h, d_g, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    HOST_TRY(cudaStreamSynchronize(st));
  }
#undef HOST_TRY
  cleanup();
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
  if (!binary && !(flags & NBCTC_FLAG_LOCKSTEP) && pipe_supported(T, B, C, Lmax)) f = std::max(f, pipe_workspace_bytes(T, B, C, Lmax));
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
