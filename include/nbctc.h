/*
 * nbctc.h -- C ABI of the B200-native no-blank CTC loss library (libnbctc.so).
 *
 * This is the drop-in boundary for the ONE hot path this repository accelerates: the
 * forward/backward dynamic programme of gotaku6629/CTC's no-blank CTC losses.  The
 * reference has no native code and no FFI of its own (it is pure PyTorch); each entry
 * point below names the reference Python interface it replaces (file:line under the
 * reference tree) -- that is what a maintainer binds it to.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in
 *     `_host`; the library never allocates, frees or retains caller memory (the `_host`
 *     convenience entry points are the one exception: they own temporary device buffers
 *     for the duration of the call);
 *   - `stream` is a `cudaStream_t` passed as `void*`; all work is enqueued on it and no
 *     entry point synchronises the host (except the `_host` ones);
 *   - return value: 0 on success, negative NBCTC_ERR_* otherwise; a human-readable
 *     message for the calling thread is available from nbctc_last_error();
 *   - layouts: logits/grad (T,B,C) row-major contiguous float32; labels (B,Lmax) int32,
 *     slots >= target_length[b] are never dereferenced (the reference pads them with -1);
 *     multi-hot targets (B,Lmax,C) float32; lengths (B) int64.
 *   - parity domain: 1 <= target_length[b] <= input_length[b] <= T and labels in [0,C).
 *     Sequences outside it get loss = +inf and an all-zero gradient (the reference
 *     returns ~1e13 and meaningless gradients there).
 */
#ifndef NBCTC_H_
#define NBCTC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBCTC_VERSION 100 /* major*10000 + minor*100 + patch */

#define NBCTC_OK 0
#define NBCTC_ERR_INVALID_ARG (-1)
#define NBCTC_ERR_UNSUPPORTED (-2)
#define NBCTC_ERR_WORKSPACE (-3)
#define NBCTC_ERR_CUDA (-4)

/* flags */
#define NBCTC_FLAG_DEFAULT 0u
#define NBCTC_FLAG_GENERIC 1u     /* force the unfused three-kernel path (debug / cross-check) */
#define NBCTC_FLAG_NO_GRAD 2u     /* loss only (validate(), train.py:486,576 runs under no_grad) */
#define NBCTC_FLAG_ALIGNED16 4u   /* caller guarantees 16-byte aligned logits and grad_logits: the workspace query
                                     then omits the room for the unaligned-tensor fallback (the call fails with
                                     NBCTC_ERR_INVALID_ARG if the promise is broken) */

#define NBCTC_FLAG_LOCKSTEP 8u    /* single-label variant: the lock-step fused kernel (stream_kernel.cuh; the default for small
                                     batches).  Without a path flag the NBCTC_PATH environment variable ("seqwarp" /
                                     "lockstep") decides, else the batch size.  (16u was round 2's pipeline kernel: removed) */
#define NBCTC_FLAG_SEQWARP 32u    /* single-label variant: the sequence-per-warp kernels (seqwarp_kernel.cuh: C <= 256, Lmax <= 64;
                                     seqwide_kernel.cuh: C % 4 == 0, C <= 1024, Lmax <= 256, 16-byte aligned tensors).
                                     Without a path flag they take the batches that fill the GPU with one sequence per
                                     warp (>= 2304 / 1536 / 384 sequences); smaller batches stay on the lock-step kernel */

#define NBCTC_FLAG_SUM_WEIGHTED 64u /* loss_sum receives weight_scalar * sum_b seq_weights[b] * loss[b] in float64 (loss_reduced
                                     without the rounding to float32) instead of the plain sum: the scalar a sharded run
                                     all-reduces, with no extra kernel between the loss call and the collective */

typedef void* nbctc_stream_t; /* cudaStream_t */

/* Library version (NBCTC_VERSION of the build). */
int nbctc_version(void);

/* Last error message of the calling thread ("" if none). Never NULL. */
const char* nbctc_last_error(void);

/* Bytes of scratch the caller must provide for one call with these shapes (same `flags` as the call).
 * `binary` = 0 for nbctc_*, 1 for nbbctc_*.  Returns 0 on invalid shapes.  The workspace pointer handed to the loss
 * calls must be 256-byte aligned (cudaMalloc and torch allocations are); NBCTC_ERR_INVALID_ARG otherwise. */
size_t nbctc_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax, int binary, uint32_t flags);

/*
 * NoBlankCTC forward (+ gradient).  Replaces NoBlankCTC.forward (NoBlankCTC.py:129-141:
 * log_softmax :136, calc_trans :90-126, computes_transition :71-87, read-out :58-68,:139)
 * and the autograd graph behind it (train.py:444 Loss.backward()).
 *
 *   loss_per_seq[b] = -log p(l_b | x_b)                                  (b < B)
 *   grad_logits[t,b,c] = w_b * (softmax(x[t,b,:])[c] - sum_{s: l_s = c} gamma_t(s))   for t < input_length[b]
 *                      = 0                                                            otherwise
 *   w_b = weight_scalar * (seq_weights ? seq_weights[b] : 1)
 * The reference's mean over the batch (NoBlankCTC.py:140) is weight_scalar = 1/B.
 * grad_logits may be NULL (or NBCTC_FLAG_NO_GRAD set) to skip the backward half.
 * loss_sum (nullable) receives sum_b loss_per_seq[b] accumulated in float64 in a fixed
 * order (bit-reproducible; this is the scalar that is all-reduced across GPUs).
 * loss_reduced (nullable) receives (float)(weight_scalar * loss_sum), i.e. the reference's
 * return value torch.mean(loss) (NoBlankCTC.py:140) when weight_scalar = 1/B.
 * Any finite float32 logits are inside the parity domain: the fused kernels work in the linear domain and hand the
 * sequences whose emissions leave the float32 range (a label more than 83 nats under its row's maximum) to a
 * log-domain kernel launched behind them on the same stream (results as for every other sequence, within 1e-5).
 */
int nbctc_loss_grad_f32(const float* logits, int64_t T, int64_t B, int64_t C,
                        const int32_t* labels, int64_t Lmax,
                        const int64_t* input_lengths, const int64_t* target_lengths,
                        float* loss_per_seq, double* loss_sum, float* loss_reduced,
                        float* grad_logits, const float* seq_weights, float weight_scalar,
                        void* workspace, size_t workspace_bytes, uint32_t flags,
                        nbctc_stream_t stream);

/*
 * NoBlankBinaryCTC forward (+ gradient).  Replaces NoBlankBinaryCTC.forward
 * (NoBlankBinaryCTC.py:139-151: sigmoid :146, BCELoss emissions :109-112/:85-88,
 * transition :72-95, read-out :58-68,:149).
 *   e[t,b,s] = (1/C) sum_c [ y log sigmoid(x) + (1-y) log(1-sigmoid(x)) ]
 *   grad_logits[t,b,c] = w_b * (sigmoid(x[t,b,c]) - sum_s gamma_t(s) y[b,s,c]) / C   for t < input_length[b]
 * targets: (B,Lmax,C) float32 in [0,1]; rows >= target_length[b] are ignored.
 */
int nbbctc_loss_grad_f32(const float* logits, int64_t T, int64_t B, int64_t C,
                         const float* targets, int64_t Lmax,
                         const int64_t* input_lengths, const int64_t* target_lengths,
                         float* loss_per_seq, double* loss_sum, float* loss_reduced,
                         float* grad_logits, const float* seq_weights, float weight_scalar,
                         void* workspace, size_t workspace_bytes, uint32_t flags,
                         nbctc_stream_t stream);

/*
 * nbctc_loss_grad_f32 with the row log-partitions exposed (the producer of the logits -- the
 * reference's LSTM head, LSTM.py:39-51, train.py:417-427 -- can fuse log-softmax's row reduction into its own
 * epilogue).  row_lse (T,B) float32 = log sum_c exp(logits[t,b,c]) (NoBlankCTC.py:136 is logits - row_lse):
 *   row_lse_in  != NULL  the loss trusts it and phase 1 reads only the label entries of every row (about half the
 *                        sectors of the logits instead of all of them);
 *   row_lse_out != NULL  receives the log-partitions this call computed, rows t < input_lengths[b] only (e.g. for the
 *                        next call on the same logits, or for a log-softmax consumer).
 * Both may be NULL (then it is nbctc_loss_grad_f32 on the sequence-per-warp kernel).  Shapes outside that kernel's
 * range (C > 256 or Lmax > 64) return NBCTC_ERR_UNSUPPORTED.
 */
int nbctc_loss_grad_lse_f32(const float* logits, int64_t T, int64_t B, int64_t C, const int32_t* labels, int64_t Lmax,
                            const int64_t* input_lengths, const int64_t* target_lengths, const float* row_lse_in,
                            float* row_lse_out, float* loss_per_seq, double* loss_sum, float* loss_reduced,
                            float* grad_logits, const float* seq_weights, float weight_scalar, void* workspace,
                            size_t workspace_bytes, uint32_t flags, nbctc_stream_t stream);

/*
 * Auxiliary cross-entropy on one frame per sequence.  The reference mixes its CTC loss with a
 * cross-entropy weighted by --alpha (opts.py:74, main.py:42, train.py:353) on the scores of one frame per sequence
 * (train.py:434: v_output[temporal-1]).  Exactly one of the two target forms is given:
 *   class_index (B) int32     nn.CrossEntropyLoss (models/__init__.py:85):  ce_b = logsumexp(x) - x[y_b]
 *   multi_hot (B,C) float32   the reference's CrossEntropy module (CrossEntropy.py:17-32): q = softmax(x),
 *                             ce_b = log sum_c exp(q_c) - sum_{n: multi_hot[b][n] == 1} q_n
 * x = logits[t_b, b, :], t_b = frame_index[b], or input_lengths[b]-1 when frame_index is NULL.
 * ce_per_seq[b] receives ce_b; when grad_logits != NULL, alpha_weight * seq_weights[b] * d ce_b / d x is ADDED to the
 * row (t_b, b) of grad_logits -- call it on the stream of, and after, *_loss_grad_f32, with
 * alpha_weight = alpha * weight_scalar (the reference averages both losses over the batch): the sum is the gradient
 * of CTC + alpha * CE.  B rows of the T*B are touched.  Sequences with t_b outside [0,T) or a class index outside
 * [0,C) get ce = +inf and no gradient.
 */
int nbctc_aux_ce_f32(const float* logits, int64_t T, int64_t B, int64_t C, const int64_t* frame_index,
                     const int64_t* input_lengths, const int32_t* class_index, const float* multi_hot, float alpha_weight,
                     const float* seq_weights, float* ce_per_seq, float* grad_logits, nbctc_stream_t stream);

/*
 * Backward-time rescale of a gradient produced by *_loss_grad_f32 with the upstream
 * gradient that autograd hands to backward() (train.py:444).  grad[t,b,c] *= g where
 * g = grad_out[0] (per_seq == 0) or grad_out[b] (per_seq != 0).  When per_seq == 0 and
 * grad_out[0] == 1.0f the kernel exits without touching memory (decided on the device,
 * no host sync).
 */
int nbctc_scale_grad_f32(float* grad_logits, int64_t T, int64_t B, int64_t C,
                         const float* grad_out, int per_seq, nbctc_stream_t stream);

/*
 * Best (Viterbi) monotone alignment on the same lattice and per-frame argmax class
 * (the reference has no implementation -- the intended use is
 * imgs/ctc_action.png and the top-k on logits at train.py:41-56,434).
 *   states[b,t] = state index of the max-score path at frame t (t < input_length[b]), -1 otherwise
 *   score[b]    = sum_t x[t,b,label[b,states[b,t]]] in float64 (nullable)
 *   argmax[t,b] = argmax_c x[t,b,c], ties -> lowest c (nullable)
 * Scores use the raw logits in float64 (the row log-partition is common to all paths),
 * ties prefer staying in the state; results are bit-exact against the float64 oracle.
 */
int nbctc_best_path_i32(const float* logits, int64_t T, int64_t B, int64_t C,
                        const int32_t* labels, int64_t Lmax,
                        const int64_t* input_lengths, const int64_t* target_lengths,
                        int32_t* states, double* score, int32_t* argmax,
                        void* workspace, size_t workspace_bytes, nbctc_stream_t stream);

/* Workspace for nbctc_best_path_i32. */
size_t nbctc_best_path_workspace_bytes(int64_t T, int64_t B, int64_t C, int64_t Lmax);

/*
 * Evaluation metrics of the reference's training loop; integer results, ties in the
 * top-k go to the LOWER class index (torch.topk leaves the order of ties open), NaN scores rank first.
 *
 * nbctc_frame_topk_i32: top-K classes of every score row, best first (`output.topk(maxk, 1, True, True)`,
 * train.py:48,65,90,118).  Row (o, i), o < n_outer, i < n_inner, starts at scores + o*stride_outer +
 * i*stride_inner (strides in elements; the C scores of a row are contiguous); pred is (n_outer, n_inner, K)
 * int32, -1 where K > C.  1 <= K <= 8.  For `output[b]` slices of a (T,B,C) tensor: n_outer = B,
 * n_inner = T, stride_outer = C, stride_inner = B*C.
 */
int nbctc_frame_topk_i32(const float* scores, int64_t n_outer, int64_t n_inner, int64_t stride_outer,
                         int64_t stride_inner, int64_t C, int K, int32_t* pred, nbctc_stream_t stream);

/*
 * Greedy monotone matching of per-frame predictions against a multi-hot transcript, for B samples at once.
 *   pred (B, frames, K) int32 from nbctc_frame_topk_i32; target (B, Lt, C) float multi-hot (> 0.5 = set);
 *   time (B) int32 = the `time` / `trans` argument per sample (rows of target in use; NULL = Lt).
 *   mode 0 = accuracy_time (train.py:111-136): correct is (B, K, frames), flag per frame;
 *   mode 1 = recall_time   (train.py:82-107):  correct is (B, K, Lt), flag per transcript row; as in the
 *            reference only the first time[b] frames are scanned (train.py:96).
 *   counts (B, K) int32 (nullable) = sum of the flags of each rank: the reference's res[k] is
 *   100 * sum(counts[b, :k]) / frames (mode 0) or / time[b] (mode 1).
 */
int nbctc_match_time_i32(const int32_t* pred, const float* target, const int32_t* time, int64_t B,
                         int64_t frames, int K, int64_t Lt, int64_t C, int mode, int32_t* correct,
                         int32_t* counts, nbctc_stream_t stream);

/*
 * Per-sample hit flags: accuracy_s (train.py:41-56; label (B) int32 class index, target NULL) or accuracy
 * (train.py:59-78; target (B, C) float multi-hot, label NULL).  pred (B, K); correct (K, B) int32.
 */
int nbctc_match_frame_i32(const int32_t* pred, const int32_t* label, const float* target, int64_t B, int K,
                          int64_t C, int32_t* correct, nbctc_stream_t stream);

/*
 * Host-buffer plugin entry points (every pointer is a HOST pointer): what a caller without device memory of its own
 * binds to.  The batch travels in chunks of sequences through three device slots -- chunk c+1 is copied in while
 * chunk c runs and the gradient of chunk c-1 is copied back -- so the call runs at the speed of the PCIe link
 * (pinned host memory: cudaHostAlloc / cudaHostRegister; pageable memory works but serialises the copies).
 * The loss (and the gradient when grad_logits_host != NULL) is complete when the call returns.  loss_sum /
 * loss_reduced are reduced on the host in float64 in ascending sequence order.  Device buffers and streams are kept
 * per device between calls; nbctc_host_release(device) frees them.  Calls for one device are serialised.
 */
int nbctc_host_release(int device);

int nbctc_loss_grad_host_f32(int device, const float* logits_host, int64_t T, int64_t B, int64_t C,
                             const int32_t* labels_host, int64_t Lmax,
                             const int64_t* input_lengths_host, const int64_t* target_lengths_host,
                             float* loss_per_seq_host, double* loss_sum_host, float* loss_reduced_host,
                             float* grad_logits_host, float weight_scalar, uint32_t flags);

int nbbctc_loss_grad_host_f32(int device, const float* logits_host, int64_t T, int64_t B, int64_t C,
                              const float* targets_host, int64_t Lmax,
                              const int64_t* input_lengths_host, const int64_t* target_lengths_host,
                              float* loss_per_seq_host, double* loss_sum_host, float* loss_reduced_host,
                              float* grad_logits_host, float weight_scalar, uint32_t flags);

/* Number of kernels this library has launched in the calling process (diagnostics / bench). */
uint64_t nbctc_kernel_launch_count(void);

/*
 * Development hook, not part of the operator interface: device buffer for the per-role cycle counters of the lock-step
 * kernel.  Only builds with -DNBCTC_PROF write to it (tools/prof_roles.py); the product build ignores the pointer.
 */
int nbctc_debug_set_prof(long long* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* NBCTC_H_ */
