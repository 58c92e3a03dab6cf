// Test-time wrapper around the head for sm_100a.
//
//   N1  RoI projection / horizontal flip, feature-RoI de-duplication (hash, unique, inverse index),
//       inverse scatter of the per-RoI scores into cls_prob [R, C+1] and test-time-augmentation averaging:
//       detectron/core/test_wsl.py:100-178 (im_detect_bbox), :181-281 (im_detect_bbox_aug),
//       :998-1059 (_get_rois_blob / _project_im_rois / _get_blobs), detectron/utils/boxes.py:246-251.
//   N2  per-class score threshold + greedy NMS + detections-per-image limit:
//       detectron/core/test_wsl.py:803-863 (box_results_with_nms_and_limit),
//       detectron/utils/cython_nms.pyx:38-93 (float32 arithmetic, '>=' on the overlap).
//   N4  MinEntropyLoss / MinEntropyLossGradient: detectron/ops/min_entropy_loss_op.cu:34-66,70-152.
//
// All of this is integer / byte / short-vector work: the kernels are sized for latency (one CTA per
// class or per image, shared-memory bitonic sorts), not for the tensor cores.
#include <algorithm>
#include <cfloat>
#include "common.cuh"

namespace nawsod {
namespace {

constexpr int kPostThreads = 1024;
constexpr int kMaxSortRois = 16384;   // NMS: 16384 * (8 + 1) B of sort keys + flags fit the 227 KB of one SM
constexpr int kMaxDedupRois = 16384;  // dedup: 16384 * (8 + 4) B (hash, row) = 192 KB; the shipped TEST.PROPOSAL_LIMIT is 9999 (configs/flickr_voc/na_wsddn_V-16-C5_1x.yaml:39)

__host__ __device__ inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// -------------------------------------------------------------------------------------------------
// N1a: rois = [batch_idx, boxes * im_scale] with an optional horizontal flip of the boxes first.
// core/test_wsl.py:1014-1027: im_rois.astype(np.float) * scale (double), hstack with the level column,
// astype(float32).  utils/boxes.py:246-251: x1' = W - x2 - 1, x2' = W - x1 - 1.
// -------------------------------------------------------------------------------------------------
__global__ void project_rois_kernel(const float* __restrict__ boxes, int R, double im_scale, double flip_width,
                                    float batch_idx, float* __restrict__ rois, const float* __restrict__ obn_in,
                                    float* __restrict__ obn_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  if (obn_in) obn_out[r] = __fadd_rn(obn_in[r], 1.0f);      // blobs['obn_scores'] = np.add(obn_scores, 1.0), core/test_wsl.py:1058
  double x1 = boxes[r * 4 + 0], y1 = boxes[r * 4 + 1], x2 = boxes[r * 4 + 2], y2 = boxes[r * 4 + 3];
  if (flip_width >= 0.0) {
    const double nx1 = flip_width - x2 - 1.0, nx2 = flip_width - x1 - 1.0;
    x1 = nx1; x2 = nx2;
  }
  rois[r * 5 + 0] = batch_idx;
  rois[r * 5 + 1] = static_cast<float>(__dmul_rn(x1, im_scale));
  rois[r * 5 + 2] = static_cast<float>(__dmul_rn(y1, im_scale));
  rois[r * 5 + 3] = static_cast<float>(__dmul_rn(x2, im_scale));
  rois[r * 5 + 4] = static_cast<float>(__dmul_rn(y2, im_scale));
}

// -------------------------------------------------------------------------------------------------
// Shared-memory bitonic sort of (key, payload) pairs, ascending by (key, payload).  P is a power of
// two; every thread of the CTA takes part.
// -------------------------------------------------------------------------------------------------
template <typename K>
__device__ __forceinline__ bool pair_less(K ka, int pa, K kb, int pb) { return ka < kb || (ka == kb && pa < pb); }

template <typename K>
__device__ void bitonic_sort_pairs(K* key, int* pay, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const K ka = key[i], kb = key[ixj];
          const int pa = pay[i], pb = pay[ixj];
          const bool up = (i & k) == 0;
          const bool swap = up ? pair_less(kb, pb, ka, pa) : pair_less(ka, pa, kb, pb);
          if (swap) { key[i] = kb; key[ixj] = ka; pay[i] = pb; pay[ixj] = pa; }
        }
      }
      __syncthreads();
    }
  }
}

// Exclusive block scan of one int per element over n elements in shared memory (in place), returns the
// total.  Simple three-phase scan: per-thread serial chunks, one warp-free Hillis-Steele over the chunk
// sums, then the offsets are added back.
__device__ int block_exclusive_scan(int* v, int n, int* chunk_sums /* [blockDim.x] */) {
  const int T = blockDim.x, t = threadIdx.x;
  const int per = (n + T - 1) / T;
  const int b = min(n, t * per), e = min(n, b + per);
  int s = 0;
  for (int i = b; i < e; ++i) { const int x = v[i]; v[i] = s; s += x; }
  chunk_sums[t] = s;
  __syncthreads();
  for (int off = 1; off < T; off <<= 1) {
    const int add = t >= off ? chunk_sums[t - off] : 0;
    __syncthreads();
    chunk_sums[t] += add;
    __syncthreads();
  }
  const int base = t ? chunk_sums[t - 1] : 0;
  const int total = chunk_sums[T - 1];
  for (int i = b; i < e; ++i) v[i] += base;
  __syncthreads();
  return total;
}

// -------------------------------------------------------------------------------------------------
// N1b: de-duplicate feature RoIs.  core/test_wsl.py:125-133:
//     v = [1, 1e3, 1e6, 1e9, 1e12]; hashes = np.round(rois * DEDUP_BOXES).dot(v)
//     _, index, inv_index = np.unique(hashes, return_index=True, return_inverse=True)
// rois * DEDUP_BOXES is a float32 product (array float32 x Python float), np.round is rint (half to
// even), the dot product runs in float64 on integer-valued terms and is therefore exact: the hash is an
// exact integer and is carried here as int64.  np.unique orders the unique hashes ascending and reports
// the FIRST occurrence of each -> sort by (hash, original index), run heads are the unique elements.
// One CTA; index [R] (entries >= num_unique repeat index[0] so a caller that skips the host round trip
// can still gather R valid rows), inv_index [R], num_unique [1].
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPostThreads, 1)
dedup_rois_kernel(const float* __restrict__ rois, int R, float dedup_scale, int32_t* __restrict__ index,
                  int32_t* __restrict__ inv_index, int32_t* __restrict__ num_unique, int32_t* __restrict__ roi_offsets) {
  extern __shared__ __align__(16) unsigned char post_smem[];
  const int P = next_pow2(R);
  long long* key = reinterpret_cast<long long*>(post_smem);
  int* pay = reinterpret_cast<int*>(post_smem + (size_t)P * sizeof(long long));
  __shared__ int chunk_sums[kPostThreads];
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    long long h = 0x7FFFFFFFFFFFFFFFll;
    if (i < R) {
      const double v[5] = {1.0, 1e3, 1e6, 1e9, 1e12};
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 5; ++k) acc = __dadd_rn(acc, __dmul_rn((double)rintf(__fmul_rn(rois[i * 5 + k], dedup_scale)), v[k]));
      h = __double2ll_rn(acc);
    }
    key[i] = h;
    pay[i] = i;
  }
  __syncthreads();
  bitonic_sort_pairs<long long>(key, pay, P);
  // run heads -> ranks, without a rank array (shared memory holds the 12-byte pairs of up to 16384 RoIs and nothing else):
  // every thread owns a contiguous chunk of the sorted sequence, counts the run heads in it, the chunk counts are scanned,
  // and a second walk over the chunk hands out the ranks.
  const int T = blockDim.x, t = threadIdx.x;
  const int per = (R + T - 1) / T;
  const int cb = min(R, t * per), ce = min(R, cb + per);
  int heads = 0;
  for (int i = cb; i < ce; ++i) heads += (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
  chunk_sums[t] = heads;
  __syncthreads();
  for (int off = 1; off < T; off <<= 1) {
    const int add = t >= off ? chunk_sums[t - off] : 0;
    __syncthreads();
    chunk_sums[t] += add;
    __syncthreads();
  }
  const int total = chunk_sums[T - 1];
  int u = t ? chunk_sums[t - 1] : 0;             // number of run heads before this chunk
  for (int i = cb; i < ce; ++i) {
    const bool head = (i == 0 || key[i] != key[i - 1]);
    if (head) index[u] = pay[i];                 // first occurrence: (hash, original index) sorts it to the run's front
    inv_index[pay[i]] = head ? u : u - 1;        // rank of the run this element belongs to
    u += head ? 1 : 0;
  }
  __syncthreads();
  const int first = pay[0];
  for (int i = total + threadIdx.x; i < R; i += blockDim.x) index[i] = first;
  if (threadIdx.x == 0) {
    num_unique[0] = total;
    if (roi_offsets) { roi_offsets[0] = 0; roi_offsets[1] = total; }   // the one-image row range of the unique set
  }
}

// rows gather: dst[i, :] = src[index[i], :] (float rows of `cols` elements)
__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ index, int n, int cols,
                                   float* __restrict__ dst) {
  const long long total = (long long)n * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i - (long long)r * cols);
    dst[i] = src[(long long)index[r] * cols + c];
  }
}

// -------------------------------------------------------------------------------------------------
// N1c: scores back to the original boxes + TTA accumulation.
//   cls_prob = concat(rois_pred[:, :1], rois_pred)            modeling/wsl_heads.py:57-67
//   scores = cls_prob[inv_index, :]                           core/test_wsl.py:173-176
//   scores_c = np.mean(scores_ts, axis=0)                     core/test_wsl.py:262-263
// np.mean over the leading axis adds the T slices one after the other in float32 and divides by T at the
// end; mode 0 assigns (first pass), mode 1 adds (later passes), nawsod_scores_finalize divides.
// -------------------------------------------------------------------------------------------------
__global__ void scatter_scores_kernel(const float* __restrict__ rois_pred, long long ld, const int32_t* __restrict__ inv_index,
                                      int R, int C, int mode, float* __restrict__ out) {
  const int K1 = C + 1;
  const long long total = (long long)R * K1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = static_cast<int>(i / K1), k = static_cast<int>(i - (long long)r * K1);
    const int u = inv_index ? inv_index[r] : r;
    const float v = rois_pred[(long long)u * ld + (k == 0 ? 0 : k - 1)];
    out[i] = mode ? __fadd_rn(out[i], v) : v;
  }
}

__global__ void scores_finalize_kernel(float* __restrict__ acc, long long n, float count) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc[i] = __fdiv_rn(acc[i], count);
}

// -------------------------------------------------------------------------------------------------
// N2: per-class threshold + greedy NMS.  One CTA per foreground class j (class 0 is background and is
// skipped, core/test_wsl.py:821-823).  Candidates are the rows with scores[:, j] > score_thresh; they
// are visited in descending score order (utils/cython_nms.pyx:45 `scores.argsort()[::-1]`; equal
// scores: higher row first, which is what reversing a stable ascending sort gives -- NumPy's own
// introsort leaves the order of ties unspecified).  The overlap arithmetic is float32 operation by
// operation as in cython_nms.pyx:44,75-86, and a box is suppressed when ovr >= thresh.
// keep [K1, R] uint8 (row 0 stays zero).
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_order_bits(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float order_bits_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void __launch_bounds__(kPostThreads, 1)
nms_per_class_kernel(const float* __restrict__ scores, const float* __restrict__ boxes, int R, int K1, float score_thresh,
                     float nms_thresh, uint8_t* __restrict__ keep, int32_t* __restrict__ num_keep) {
  extern __shared__ __align__(16) unsigned char post_smem[];
  const int j = blockIdx.x + 1;
  const int P = next_pow2(R);
  unsigned long long* key = reinterpret_cast<unsigned long long*>(post_smem);
  uint8_t* sup = post_smem + (size_t)P * sizeof(unsigned long long);
  __shared__ int n_cand_s, n_keep_s;
  if (threadIdx.x == 0) { n_cand_s = 0; n_keep_s = 0; }
  __syncthreads();
  // descending (score, row): sort ascending on the complemented key so candidates come first
  int local = 0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    unsigned long long k = ~0ull;                                     // non-candidates sort to the end
    if (i < R) {
      const float s = scores[(size_t)i * K1 + j];
      if (s > score_thresh) {
        k = ~((static_cast<unsigned long long>(float_order_bits(s)) << 32) | static_cast<unsigned int>(i));
        ++local;
      }
    }
    key[i] = k;
    sup[i] = 0;
  }
  if (local) atomicAdd(&n_cand_s, local);
  __syncthreads();
  // bitonic sort of the keys alone (the row index is the low half of the key)
  for (int k = 2; k <= P; k <<= 1) {
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int ixj = i ^ jj;
        if (ixj > i) {
          const unsigned long long a = key[i], b = key[ixj];
          const bool up = (i & k) == 0;
          if (up ? (b < a) : (a < b)) { key[i] = b; key[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  const int n = n_cand_s;
  const float4* box4 = reinterpret_cast<const float4*>(boxes);
  for (int _i = 0; _i < n; ++_i) {
    if (sup[_i]) continue;                                            // uniform: sup[_i] is final here
    const int i = static_cast<int>(~key[_i] & 0xFFFFFFFFull);
    const float4 bi = __ldg(box4 + i);
    const float iarea = __fmul_rn(__fadd_rn(__fsub_rn(bi.z, bi.x), 1.f), __fadd_rn(__fsub_rn(bi.w, bi.y), 1.f));
    for (int _j = _i + 1 + threadIdx.x; _j < n; _j += blockDim.x) {
      if (sup[_j]) continue;
      const int jr = static_cast<int>(~key[_j] & 0xFFFFFFFFull);
      const float4 bj = __ldg(box4 + jr);
      const float jarea = __fmul_rn(__fadd_rn(__fsub_rn(bj.z, bj.x), 1.f), __fadd_rn(__fsub_rn(bj.w, bj.y), 1.f));
      const float xx1 = bi.x >= bj.x ? bi.x : bj.x, yy1 = bi.y >= bj.y ? bi.y : bj.y;
      const float xx2 = bi.z <= bj.z ? bi.z : bj.z, yy2 = bi.w <= bj.w ? bi.w : bj.w;
      const float w0 = __fadd_rn(__fsub_rn(xx2, xx1), 1.f), h0 = __fadd_rn(__fsub_rn(yy2, yy1), 1.f);
      const float w = 0.0f >= w0 ? 0.0f : w0, h = 0.0f >= h0 ? 0.0f : h0;
      const float inter = __fmul_rn(w, h);
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
      if (ovr >= nms_thresh) sup[_j] = 1;
    }
    __syncthreads();
  }
  __syncthreads();
  int kept = 0;
  for (int i = threadIdx.x; i < R; i += blockDim.x) keep[(size_t)j * R + i] = 0;
  __syncthreads();
  for (int _i = threadIdx.x; _i < n; _i += blockDim.x)
    if (!sup[_i]) { keep[(size_t)j * R + static_cast<int>(~key[_i] & 0xFFFFFFFFull)] = 1; ++kept; }
  if (kept) atomicAdd(&n_keep_s, kept);
  __syncthreads();
  if (threadIdx.x == 0) num_keep[j] = n_keep_s;
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < R; i += blockDim.x) keep[i] = 0;
    if (threadIdx.x == 0) num_keep[0] = 0;
  }
}

// Detections-per-image limit (core/test_wsl.py:852-860): if more than `limit` detections survive NMS
// over all classes, image_thresh = np.sort(image_scores)[-limit] and only scores >= image_thresh stay.
// One CTA: 4-pass radix select (8 bits per pass, from the top) of the limit-th largest kept score.
__global__ void __launch_bounds__(kPostThreads, 1)
limit_detections_kernel(const float* __restrict__ scores, int R, int K1, int limit, uint8_t* __restrict__ keep,
                        int32_t* __restrict__ num_keep, float* __restrict__ image_thresh) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int prefix_s, want_s, total_s;
  const long long N = (long long)K1 * R;
  if (threadIdx.x == 0) total_s = 0;
  __syncthreads();
  unsigned int local = 0;
  for (long long i = threadIdx.x; i < N; i += blockDim.x) local += keep[i];
  atomicAdd(&total_s, local);
  __syncthreads();
  if (image_thresh && threadIdx.x == 0) image_thresh[0] = -FLT_MAX;
  if (limit <= 0 || total_s <= (unsigned int)limit) return;
  if (threadIdx.x == 0) { prefix_s = 0; want_s = (unsigned int)limit; }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    const unsigned int prefix = prefix_s;
    const unsigned int mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (long long i = threadIdx.x; i < N; i += blockDim.x) {
      if (!keep[i]) continue;
      const int j = static_cast<int>(i / R), r = static_cast<int>(i - (long long)j * R);
      const unsigned int u = float_order_bits(scores[(size_t)r * K1 + j]);
      if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int want = want_s, b = 255;
      for (;; --b) {                     // walk down from the top bucket until the want-th largest is inside
        if (hist[b] >= want) break;
        want -= hist[b];
        if (b == 0) break;
      }
      prefix_s = prefix | (b << shift);
      want_s = want;
    }
    __syncthreads();
  }
  const unsigned int tbits = prefix_s;
  if (image_thresh && threadIdx.x == 0) image_thresh[0] = order_bits_float(tbits);
  for (int j = threadIdx.x; j < K1; j += blockDim.x) num_keep[j] = 0;
  __syncthreads();
  for (long long i = threadIdx.x; i < N; i += blockDim.x) {
    if (!keep[i]) continue;
    const int j = static_cast<int>(i / R), r = static_cast<int>(i - (long long)j * R);
    if (float_order_bits(scores[(size_t)r * K1 + j]) >= tbits) atomicAdd(&num_keep[j], 1);
    else keep[i] = 0;
  }
}

// -------------------------------------------------------------------------------------------------
// N4: MinEntropyLoss([X, L] -> Y) and its gradient (ops/min_entropy_loss_op.cu:34-66, 70-152).
//   fwd: over (n, c) with L[0, c] >= 0.5: loss += -p log p, p = max(X[n, c], 1e-20); Y = loss / (1 + count)
//   bwd: dX[n, c] = min(dY / (1 + count) * (-1 - log p), 1e4) on the same elements, 0 elsewhere.
// The reference sums with float atomics in launch order; here per-thread float partials are combined in
// a fixed tree (deterministic); the results agree to float rounding of the sum.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPostThreads, 1)
min_entropy_fwd_kernel(const float* __restrict__ X, const float* __restrict__ L, int N, int C, float* __restrict__ Y,
                       float* __restrict__ norm_out) {
  __shared__ double red[kPostThreads];
  __shared__ int cnt[kPostThreads];
  const long long total = (long long)N * C;
  float s = 0.f;
  int k = 0;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    const int c = static_cast<int>(i % C);
    if (L[c] < 0.5f) continue;
    const float prob = fmaxf(X[i], 1e-20f);
    s += -prob * logf(prob);
    ++k;
  }
  red[threadIdx.x] = (double)s;
  cnt[threadIdx.x] = k;
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    if (threadIdx.x < off) { red[threadIdx.x] += red[threadIdx.x + off]; cnt[threadIdx.x] += cnt[threadIdx.x + off]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = 1.f + static_cast<float>(cnt[0]);
    Y[0] = __fdiv_rn(static_cast<float>(red[0]), norm);
    if (norm_out) norm_out[0] = norm;
  }
}

__global__ void min_entropy_count_kernel(const float* __restrict__ L, int N, int C, float* __restrict__ norm) {
  // count = N * #{c : L[0, c] >= 0.5}; one warp is plenty
  int k = 0;
  for (int c = threadIdx.x; c < C; c += 32) k += (L[c] >= 0.5f) ? 1 : 0;
  for (int o = 16; o; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
  if (threadIdx.x == 0) norm[0] = 1.f + static_cast<float>((long long)k * N);
}

__global__ void min_entropy_bwd_kernel(const float* __restrict__ X, const float* __restrict__ L, const float* __restrict__ dY,
                                       const float* __restrict__ norm, int N, int C, float* __restrict__ dX) {
  const long long total = (long long)N * C;
  const float scale = __fdiv_rn(dY[0], norm[0]);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % C);
    float d = 0.f;
    if (!(L[c] < 0.5f)) {
      const float prob = fmaxf(X[i], 1e-20f);
      d = fminf(scale * (-1.f + (-1.f) * logf(prob)), 1e4f);
    }
    dX[i] = d;
  }
}

int grid_for(long long total, int threads = 256) {
  return (int)std::max<long long>(1, std::min<long long>((total + threads - 1) / threads, (long long)sm_count() * 8));
}

}  // namespace
}  // namespace nawsod

using namespace nawsod;

extern "C" int nawsod_project_rois(const float* boxes, int R, double im_scale, double flip_width, int batch_idx,
                                   float* rois, const float* obn_scores, float* obn_out, void* stream) {
  NAWSOD_REQUIRE(R >= 0, NAWSOD_ERR_SHAPE, "project_rois: negative R");
  if (R == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(boxes && rois, NAWSOD_ERR_ARG, "project_rois: null pointer");
  NAWSOD_REQUIRE((obn_scores == nullptr) == (obn_out == nullptr), NAWSOD_ERR_ARG,
                 "project_rois: obn_scores and obn_out must both be given or both be NULL");
  project_rois_kernel<<<(R + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      boxes, R, im_scale, flip_width, static_cast<float>(batch_idx), rois, obn_scores, obn_out);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_dedup_rois(const float* rois, int R, float dedup_scale, int32_t* index, int32_t* inv_index,
                                 int32_t* num_unique, int32_t* roi_offsets, void* stream) {
  NAWSOD_REQUIRE(R > 0, NAWSOD_ERR_SHAPE, "dedup_rois: need R > 0 (got %d)", R);
  NAWSOD_REQUIRE(R <= kMaxDedupRois, NAWSOD_ERR_UNSUPPORTED, "dedup_rois: R=%d > %d", R, kMaxDedupRois);
  NAWSOD_REQUIRE(rois && index && inv_index && num_unique, NAWSOD_ERR_ARG, "dedup_rois: null pointer");
  NAWSOD_REQUIRE(dedup_scale > 0.f, NAWSOD_ERR_ARG, "dedup_rois: DEDUP_BOXES must be > 0");
  const int P = next_pow2(R);
  const size_t smem = (size_t)P * (sizeof(long long) + sizeof(int));
  static bool attr_set = false;
  if (!attr_set) {
    NAWSOD_CUDA_OK(cudaFuncSetAttribute(dedup_rois_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((size_t)kMaxDedupRois * (sizeof(long long) + sizeof(int)))));
    attr_set = true;
  }
  dedup_rois_kernel<<<1, kPostThreads, smem, static_cast<cudaStream_t>(stream)>>>(rois, R, dedup_scale, index, inv_index,
                                                                                 num_unique, roi_offsets);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_gather_rows(const float* src, const int32_t* index, int n, int cols, float* dst, void* stream) {
  NAWSOD_REQUIRE(n >= 0 && cols >= 0, NAWSOD_ERR_SHAPE, "gather_rows: negative size");
  if (n == 0 || cols == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(src && index && dst, NAWSOD_ERR_ARG, "gather_rows: null pointer");
  gather_rows_kernel<<<grid_for((long long)n * cols), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, index, n, cols, dst);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_scatter_scores(const float* rois_pred, int64_t ld, const int32_t* inv_index, int R, int C,
                                     int accumulate, float* cls_prob, void* stream) {
  NAWSOD_REQUIRE(R >= 0 && C > 0 && ld >= C, NAWSOD_ERR_SHAPE, "scatter_scores: bad shape R=%d C=%d ld=%lld", R, C, (long long)ld);
  if (R == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(rois_pred && cls_prob, NAWSOD_ERR_ARG, "scatter_scores: null pointer");
  scatter_scores_kernel<<<grid_for((long long)R * (C + 1)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rois_pred, ld, inv_index, R, C, accumulate ? 1 : 0, cls_prob);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_scores_finalize(float* acc, int64_t n, int count, void* stream) {
  NAWSOD_REQUIRE(n >= 0 && count > 0, NAWSOD_ERR_SHAPE, "scores_finalize: need n >= 0 and count > 0");
  if (n == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(acc, NAWSOD_ERR_ARG, "scores_finalize: null pointer");
  scores_finalize_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(acc, n, static_cast<float>(count));
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_nms_and_limit(const float* scores, const float* boxes, int R, int num_classes, float score_thresh,
                                    float nms_thresh, int detections_per_im, uint8_t* keep, int32_t* num_keep,
                                    float* image_thresh, void* stream) {
  NAWSOD_REQUIRE(R > 0 && num_classes >= 2, NAWSOD_ERR_SHAPE, "nms_and_limit: need R > 0 and at least one foreground class");
  NAWSOD_REQUIRE(R <= kMaxSortRois, NAWSOD_ERR_UNSUPPORTED, "nms_and_limit: R=%d > %d", R, kMaxSortRois);
  NAWSOD_REQUIRE(scores && boxes && keep && num_keep, NAWSOD_ERR_ARG, "nms_and_limit: null pointer");
  NAWSOD_REQUIRE(aligned16(boxes), NAWSOD_ERR_ALIGN, "nms_and_limit: boxes must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int P = next_pow2(R);
  const size_t smem = (size_t)P * (sizeof(unsigned long long) + 1);
  static bool attr_set = false;
  if (!attr_set) {
    NAWSOD_CUDA_OK(cudaFuncSetAttribute(nms_per_class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((size_t)kMaxSortRois * (sizeof(unsigned long long) + 1))));
    attr_set = true;
  }
  nms_per_class_kernel<<<num_classes - 1, kPostThreads, smem, st>>>(scores, boxes, R, num_classes, score_thresh, nms_thresh,
                                                                    keep, num_keep);
  NAWSOD_LAUNCH_OK();
  limit_detections_kernel<<<1, kPostThreads, 0, st>>>(scores, R, num_classes, detections_per_im, keep, num_keep, image_thresh);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_min_entropy_loss_fwd(const float* X, const float* L, int N, int C, int B, float* Y, float* norm,
                                           void* stream) {
  NAWSOD_REQUIRE(N > 0 && C > 0, NAWSOD_ERR_SHAPE, "min_entropy_loss: need N, C > 0");
  NAWSOD_REQUIRE(B == 1, NAWSOD_ERR_SHAPE, "min_entropy_loss: L must have one row (got %d)", B);   // CAFFE_ENFORCE_EQ(L.dim32(0), 1)
  NAWSOD_REQUIRE(X && L && Y, NAWSOD_ERR_ARG, "min_entropy_loss: null pointer");
  min_entropy_fwd_kernel<<<1, kPostThreads, 0, static_cast<cudaStream_t>(stream)>>>(X, L, N, C, Y, norm);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_min_entropy_loss_bwd(const float* X, const float* L, const float* dY, int N, int C, int B, float* dX,
                                           float* norm_ws, void* stream) {
  NAWSOD_REQUIRE(N > 0 && C > 0, NAWSOD_ERR_SHAPE, "min_entropy_loss_grad: need N, C > 0");
  NAWSOD_REQUIRE(B == 1, NAWSOD_ERR_SHAPE, "min_entropy_loss_grad: L must have one row (got %d)", B);
  NAWSOD_REQUIRE(X && L && dY && dX && norm_ws, NAWSOD_ERR_ARG, "min_entropy_loss_grad: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  min_entropy_count_kernel<<<1, 32, 0, st>>>(L, N, C, norm_ws);
  NAWSOD_LAUNCH_OK();
  min_entropy_bwd_kernel<<<grid_for((long long)N * C), 256, 0, st>>>(X, L, dY, norm_ws, N, C, dX);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}
