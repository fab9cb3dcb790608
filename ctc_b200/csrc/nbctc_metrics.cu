// Evaluation metrics of the reference's training loop on the device: per-frame top-k and the
// greedy monotone matching of predictions against a multi-hot transcript (train.py:41-136).  Integer results,
// bit-exact against the CPU restatement used by the tests.
//
//   frame_topk   one warp per score row.  Scores become order-preserving 32-bit keys (NaN highest, as torch.topk
//                does; -0 == +0), a lane keeps the K best of its strided share in registers, K warp-wide arg-max
//                rounds merge them.  Ties go to the LOWER class index.  HBM bound: 4*C bytes read per row.
//   match_time   one thread per (sample, rank): the matching is a sequential scan (the cursor into the transcript
//                only moves forward), frames x transcript rows gathers in the worst case -- tiny next to the top-k.
//   match_frame  one thread per (sample, rank): class-index or multi-hot lookup.
#include <algorithm>

#include "common.cuh"

namespace nbctc {
namespace {

constexpr int kMaxK = 8;

__device__ __forceinline__ uint32_t score_key(float v) {
  if (v != v) return 0xFFFFFFFFu;  // NaN sorts first (torch.topk semantics)
  v += 0.0f;                       // -0 -> +0
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// key = (score key << 32) | (0xFFFFFFFF - class): a plain unsigned max prefers the higher score, then the lower class
__device__ __forceinline__ unsigned long long pack_key(float v, uint32_t c) {
  return ((unsigned long long)score_key(v) << 32) | (unsigned long long)(0xFFFFFFFFu - c);
}

template <int K>
__device__ __forceinline__ void insert_key(unsigned long long (&best)[K], unsigned long long k) {
  if (k <= best[K - 1]) return;
#pragma unroll
  for (int i = K - 1; i >= 1; --i) {
    const bool up = k > best[i - 1];
    best[i] = up ? best[i - 1] : (k > best[i] ? k : best[i]);
  }
  best[0] = k > best[0] ? k : best[0];
}

template <int K>
__global__ void __launch_bounds__(256) frame_topk_kernel(const float* __restrict__ x, int64_t n_outer, int64_t n_inner,
                                                         int64_t stride_outer, int64_t stride_inner, int C, int k_out,
                                                         int32_t* __restrict__ pred) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = n_outer * n_inner;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    const float* row = x + (r / n_inner) * stride_outer + (r % n_inner) * stride_inner;
    unsigned long long best[K];
#pragma unroll
    for (int i = 0; i < K; ++i) best[i] = 0ull;  // below every real key (a real key has class <= 2^32-2)
    // 16-byte loads when the row allows it: lane l takes chunk l, l+32, ...
    if ((C & 3) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15) == 0)) {
      const float4* r4 = reinterpret_cast<const float4*>(row);
      for (int q = lane; q < (C >> 2); q += 32) {
        const float4 v = __ldcs(r4 + q);
        insert_key<K>(best, pack_key(v.x, 4 * q));
        insert_key<K>(best, pack_key(v.y, 4 * q + 1));
        insert_key<K>(best, pack_key(v.z, 4 * q + 2));
        insert_key<K>(best, pack_key(v.w, 4 * q + 3));
      }
    } else {
      for (int c = lane; c < C; c += 32) insert_key<K>(best, pack_key(__ldcs(row + c), c));
    }
    // K rounds: warp arg-max over the heads, the winner pops
    for (int i = 0; i < k_out; ++i) {
      unsigned long long m = best[0];
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
      }
      if (best[0] == m && m != 0ull) {
#pragma unroll
        for (int j = 0; j < K - 1; ++j) best[j] = best[j + 1];
        best[K - 1] = 0ull;
      }
      if (lane == 0) pred[r * k_out + i] = (m == 0ull) ? -1 : (int32_t)(0xFFFFFFFFu - (uint32_t)(m & 0xFFFFFFFFull));
    }
  }
}

// Rows of up to 32*E classes: the whole row sits in registers as order-preserving SIGNED keys (lane l holds classes
// l, l+32, ...).  x + 0 turns -0 into +0 and any NaN into the canonical 0x7fffffff, the xor folds the negative floats:
// NaN > +inf > ... > -inf > INT_MIN (= empty slot).  One round = lane-local max, REDUX max over the warp, lowest class
// holding that key (REDUX min), clear it.  kExact: C > 32*(E-1), only the last slot needs a bounds check.
__device__ __forceinline__ int score_skey(float v) {
  const int k = __float_as_int(__fadd_rn(v, 0.0f));
  return k ^ ((k >> 31) & 0x7fffffff);
}

template <int E, int KMAX, bool kExact>
__global__ void __launch_bounds__(256) frame_topk_reg_kernel(const float* __restrict__ x, int64_t n_outer, int64_t n_inner,
                                                             int64_t stride_outer, int64_t stride_inner, int C, int k_out,
                                                             int32_t* __restrict__ pred) {
  constexpr int kEmpty = (int)0x80000000;
  const int lane = threadIdx.x & 31;
  const int64_t rows = n_outer * n_inner;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  // (outer, inner) of the row advance incrementally: one division per warp, none per row
  int64_t ro = r / n_inner, ri = r % n_inner;
  const int64_t step_o = warps / n_inner, step_i = warps % n_inner;
  for (; r < rows; r += warps, ro += step_o, ri += step_i) {
    if (ri >= n_inner) {
      ri -= n_inner;
      ++ro;
    }
    const float* row = x + ro * stride_outer + ri * stride_inner + lane;
    int key[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if ((kExact && e < E - 1) || lane + 32 * e < C)
        key[e] = score_skey(__ldcs(row + 32 * e));
      else
        key[e] = kEmpty;
    }
    int32_t mine = -1;
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      if (i >= k_out) break;
      int lm = key[0];
#pragma unroll
      for (int e = 1; e < E; ++e) lm = max(lm, key[e]);
      const int m = __reduce_max_sync(0xffffffffu, lm);
      if (m == kEmpty) break;  // fewer than k_out classes: the rest stays -1
      int le = 1 << 20;        // lowest slot of this lane holding m (sentinel: above every real slot)
#pragma unroll
      for (int e = E - 1; e >= 0; --e) le = (key[e] == m) ? e : le;
      const uint32_t c = __reduce_min_sync(0xffffffffu, (uint32_t)le * 32u + (uint32_t)lane);  // class = 32*slot + lane
      const int ce = (c & 31u) == (uint32_t)lane ? (int)(c >> 5) : -1;
#pragma unroll
      for (int e = 0; e < E; ++e) key[e] = (e == ce) ? kEmpty : key[e];
      if (lane == i) mine = (int32_t)c;
    }
    if (lane < k_out) pred[r * k_out + lane] = mine;
  }
}

// accuracy_time / recall_time for transcripts of up to 64 rows: one CTA per sample, one warp per rank.
// Step 1 turns the sample's multi-hot transcript (read once, coalesced) into one 64-bit row mask per class in shared
// memory; step 2 looks the masks of 32 frames up in parallel and walks them in order with a register cursor.
__global__ void __launch_bounds__(256) match_time_mask_kernel(const int32_t* __restrict__ pred, const float* __restrict__ target,
                                                              const int32_t* __restrict__ time, int frames, int K, int Lt, int C,
                                                              int mode, int32_t* __restrict__ correct,
                                                              int32_t* __restrict__ counts) {
  extern __shared__ unsigned long long cmask[];  // [C]
  const int b = blockIdx.x;
  int tb = time ? time[b] : Lt;
  tb = tb < 0 ? 0 : (tb > Lt ? Lt : tb);
  const float* tg = target + (int64_t)b * Lt * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    unsigned long long m = 0ull;
    for (int t = 0; t < tb; ++t) m |= (unsigned long long)(__ldcs(tg + (int64_t)t * C + c) > 0.5f) << t;
    cmask[c] = m;
  }
  __syncthreads();
  const int i = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (i >= K) return;
  const int width = mode == 0 ? frames : Lt;
  int32_t* out = correct + ((int64_t)b * K + i) * width;
  const int nj = mode == 0 ? frames : (tb < frames ? tb : frames);
  int cur = 0, n = 0;
  unsigned long long rows_hit = 0ull;
  for (int j0 = 0; j0 < nj; j0 += 32) {
    unsigned long long mk = 0ull;
    if (j0 + lane < nj) {
      const int c = pred[((int64_t)b * frames + j0 + lane) * K + i];
      if (c >= 0 && c < C) mk = cmask[c];
    }
    uint32_t flags = 0u;
    const int lim = min(32, nj - j0);
    for (int l = 0; l < lim; ++l) {
      const unsigned long long m = __shfl_sync(0xffffffffu, mk, l) >> cur;
      if (m) {
        cur += __ffsll((long long)m) - 1;
        flags |= 1u << l;
        rows_hit |= 1ull << cur;
      }
    }
    if (mode == 0) {
      if (j0 + lane < frames) out[j0 + lane] = (flags >> lane) & 1u;
      n += __popc(flags);
    }
  }
  if (mode == 1) {
    for (int t = lane; t < Lt; t += 32) out[t] = (int32_t)((rows_hit >> t) & 1ull);
    n = __popcll(rows_hit);
  }
  if (counts && lane == 0) counts[b * K + i] = n;
}

// mode 0: accuracy_time (train.py:111-136) -- correct (B,K,frames), flag per FRAME
// mode 1: recall_time   (train.py:82-107)  -- correct (B,K,Lt), flag per TRANSCRIPT ROW; only the first time[b] frames
//         are looked at (train.py:96 loops over correct.shape[1] = trans)
__global__ void match_time_kernel(const int32_t* __restrict__ pred, const float* __restrict__ target,
                                  const int32_t* __restrict__ time, int B, int frames, int K, int Lt, int C, int mode,
                                  int32_t* __restrict__ correct, int32_t* __restrict__ counts) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * K) return;
  const int b = idx / K, i = idx % K;
  int tb = time ? time[b] : Lt;
  tb = tb < 0 ? 0 : (tb > Lt ? Lt : tb);
  const int width = mode == 0 ? frames : Lt;
  int32_t* out = correct + ((int64_t)b * K + i) * width;
  const float* tg = target + (int64_t)b * Lt * C;
  const int nj = mode == 0 ? frames : (tb < frames ? tb : frames);
  for (int j = 0; j < width; ++j) out[j] = 0;
  int cur = 0, n = 0;
  for (int j = 0; j < nj; ++j) {
    const int c = pred[((int64_t)b * frames + j) * K + i];
    if (c < 0 || c >= C) continue;
    for (int t = cur; t < tb; ++t) {
      if (tg[(int64_t)t * C + c] > 0.5f) {
        const int o = mode == 0 ? j : t;
        n += 1 - out[o];  // a transcript row can be hit again by a later frame (recall): count it once
        out[o] = 1;
        cur = t;
        break;
      }
    }
  }
  if (counts) counts[idx] = n;
}

// accuracy_s (train.py:41-56, label != null) / accuracy (train.py:59-78, multi-hot target): correct (K,B)
__global__ void match_frame_kernel(const int32_t* __restrict__ pred, const int32_t* __restrict__ label,
                                   const float* __restrict__ target, int B, int K, int C, int32_t* __restrict__ correct) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * K) return;
  const int i = idx / B, b = idx % B;
  const int c = pred[(int64_t)b * K + i];
  int ok = 0;
  if (c >= 0 && c < C) ok = label ? (label[b] == c) : (target[(int64_t)b * C + c] > 0.5f);
  correct[idx] = ok;
}

}  // namespace
}  // namespace nbctc

using namespace nbctc;

extern "C" int nbctc_frame_topk_i32(const float* scores, int64_t n_outer, int64_t n_inner, int64_t stride_outer,
                                    int64_t stride_inner, int64_t C, int K, int32_t* pred, nbctc_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_outer < 0 || n_inner < 0 || C < 1 || C > 0x7fffffff || K < 1 || K > kMaxK) {
    set_error("frame_topk: bad shape (n_outer=%lld n_inner=%lld C=%lld K=%d; 1 <= K <= %d)", (long long)n_outer,
              (long long)n_inner, (long long)C, K, kMaxK);
    return NBCTC_ERR_INVALID_ARG;
  }
  const int64_t rows = n_outer * n_inner;
  if (rows == 0) return NBCTC_OK;
  if (!scores || !pred) {
    set_error("frame_topk: null pointer");
    return NBCTC_ERR_INVALID_ARG;
  }
  int dev = 0, sms = 148;
  NBCTC_CUDA_CHECK(cudaGetDevice(&dev));
  NBCTC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t want = (rows + 7) / 8;
  const int grid = (int)std::min<int64_t>(want, (int64_t)sms * 8 * 4);
#define NBCTC_TOPK_REG(E_, X_)                                                                                    \
  do {                                                                                                              \
    if (K == 1)                                                                                                     \
      frame_topk_reg_kernel<E_, 1, X_><<<grid, 256, 0, stream>>>(scores, n_outer, n_inner, stride_outer, stride_inner, \
                                                                 (int)C, K, pred);                                  \
    else if (K <= 5)                                                                                                \
      frame_topk_reg_kernel<E_, 5, X_><<<grid, 256, 0, stream>>>(scores, n_outer, n_inner, stride_outer, stride_inner, \
                                                                 (int)C, K, pred);                                  \
    else                                                                                                            \
      frame_topk_reg_kernel<E_, 8, X_><<<grid, 256, 0, stream>>>(scores, n_outer, n_inner, stride_outer, stride_inner, \
                                                                 (int)C, K, pred);                                  \
  } while (0)
  if (C <= 32) NBCTC_TOPK_REG(1, true);
  else if (C <= 64) NBCTC_TOPK_REG(2, true);
  else if (C <= 96) NBCTC_TOPK_REG(3, true);
  else if (C <= 128) NBCTC_TOPK_REG(4, true);
  else if (C <= 160) NBCTC_TOPK_REG(5, true);
  else if (C <= 192) NBCTC_TOPK_REG(6, true);
  else if (C <= 224) NBCTC_TOPK_REG(7, true);
  else if (C <= 256) NBCTC_TOPK_REG(8, true);
  else if (C <= 384) NBCTC_TOPK_REG(12, false);
  else if (C <= 512) NBCTC_TOPK_REG(16, false);
#undef NBCTC_TOPK_REG
  else if (K <= 1)
    frame_topk_kernel<1><<<grid, 256, 0, stream>>>(scores, n_outer, n_inner, stride_outer, stride_inner, (int)C, K, pred);
  else if (K <= 5)
    frame_topk_kernel<5><<<grid, 256, 0, stream>>>(scores, n_outer, n_inner, stride_outer, stride_inner, (int)C, K, pred);
  else
    frame_topk_kernel<8><<<grid, 256, 0, stream>>>(scores, n_outer, n_inner, stride_outer, stride_inner, (int)C, K, pred);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

extern "C" int nbctc_match_time_i32(const int32_t* pred, const float* target, const int32_t* time, int64_t B,
                                    int64_t frames, int K, int64_t Lt, int64_t C, int mode, int32_t* correct,
                                    int32_t* counts, nbctc_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (B < 0 || frames < 0 || Lt < 0 || C < 1 || K < 1 || K > kMaxK || (mode != 0 && mode != 1) ||
      B * K > 0x7fffffff || frames > 0x7fffffff || Lt > 0x7fffffff || C > 0x7fffffff) {
    set_error("match_time: bad argument (B=%lld frames=%lld K=%d Lt=%lld C=%lld mode=%d)", (long long)B,
              (long long)frames, K, (long long)Lt, (long long)C, mode);
    return NBCTC_ERR_INVALID_ARG;
  }
  if (B == 0) return NBCTC_OK;
  if (!correct || (frames > 0 && !pred) || (Lt > 0 && !target)) {
    set_error("match_time: null pointer");
    return NBCTC_ERR_INVALID_ARG;
  }
  if (Lt <= 64 && (size_t)C * 8 <= 200 * 1024) {
    const size_t smem = (size_t)C * 8;
    if (smem > 48 * 1024)
      NBCTC_CUDA_CHECK(cudaFuncSetAttribute(match_time_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    match_time_mask_kernel<<<(unsigned)B, 32 * std::max(K, 4), smem, stream>>>(pred, target, time, (int)frames, K, (int)Lt,
                                                                            (int)C, mode, correct, counts);
    NBCTC_LAUNCH_CHECK();
    return NBCTC_OK;
  }
  const int n = (int)(B * K);
  match_time_kernel<<<(n + 127) / 128, 128, 0, stream>>>(pred, target, time, (int)B, (int)frames, K, (int)Lt, (int)C, mode,
                                                        correct, counts);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}

extern "C" int nbctc_match_frame_i32(const int32_t* pred, const int32_t* label, const float* target, int64_t B, int K,
                                     int64_t C, int32_t* correct, nbctc_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (B < 0 || C < 1 || K < 1 || K > kMaxK || B * K > 0x7fffffff || C > 0x7fffffff || ((label != nullptr) == (target != nullptr))) {
    set_error("match_frame: bad argument (B=%lld K=%d C=%lld; exactly one of label/target)", (long long)B, K, (long long)C);
    return NBCTC_ERR_INVALID_ARG;
  }
  if (B == 0) return NBCTC_OK;
  if (!pred || !correct) {
    set_error("match_frame: null pointer");
    return NBCTC_ERR_INVALID_ARG;
  }
  const int n = (int)(B * K);
  match_frame_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pred, label, target, (int)B, K, (int)C, correct);
  NBCTC_LAUNCH_CHECK();
  return NBCTC_OK;
}
