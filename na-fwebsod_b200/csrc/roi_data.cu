// Training-input contract of the head for sm_100a: what the reference's
// loader threads do on the host with NumPy between the proposal pickle and the `rois` / `obn_scores` /
// `labels_oh` blobs the head consumes, as stream-ordered device calls (so a minibatch whose proposals
// already live in HBM never visits the host):
//   tools/convert_mcg.py:37-49                     MCG .mat boxes -> 0-indexed (x1,y1,x2,y2) uint16
//   detectron/roi_data/wsl.py:87-181  (_sample_rois)       first BATCH_SIZE_PER_IM boxes, obn_scores + 1, labels
//   detectron/roi_data/wsl.py:212-225 (_project_im_rois)   clip to the crop window, shift, scale
//   detectron/roi_data/loader_wsl.py:149-168               bagging-mixup of two images of one class
// Integer / short-vector work: one thread per box (or per 16 bytes for the image mix), nothing to tile.
#include <algorithm>
#include "common.cuh"

namespace nawsod {
namespace {

constexpr int kThreads = 256;

// np.maximum / np.minimum on a float32 column and an integer scalar: NaN in the column propagates
__device__ __forceinline__ float np_max(float a, float b) { return (a >= b || a != a) ? a : b; }
__device__ __forceinline__ float np_min(float a, float b) { return (a <= b || a != a) ? a : b; }

// wsl.py:212-225 + :109-111,160: clip (x1,y1 from below first; x2,y2 from above first), subtract the crop origin and
// scale in double (float32 array - int32 tile is float64 in NumPy), then astype(float32) behind the batch-index column.
__global__ void sample_rois_kernel(const float* __restrict__ boxes, int R, double im_scale, float cx1, float cy1, float cx2,
                                   float cy2, float batch_idx, float* __restrict__ rois, const float* __restrict__ obn_in,
                                   float* __restrict__ obn_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  if (obn_in) obn_out[r] = __fadd_rn(obn_in[r], 1.0f);          // sampled_scores = np.add(obn_scores, 1.0), wsl.py:103
  const float4 b = *reinterpret_cast<const float4*>(boxes + 4 * (size_t)r);
  const float x1 = np_min(np_max(b.x, cx1), cx2);
  const float y1 = np_min(np_max(b.y, cy1), cy2);
  const float x2 = np_max(np_min(b.z, cx2), cx1);
  const float y2 = np_max(np_min(b.w, cy2), cy1);
  const double ox = cx1, oy = cy1;
  float* o = rois + 5 * (size_t)r;
  o[0] = batch_idx;
  o[1] = static_cast<float>(__dmul_rn(__dsub_rn(x1, ox), im_scale));
  o[2] = static_cast<float>(__dmul_rn(__dsub_rn(y1, oy), im_scale));
  o[3] = static_cast<float>(__dmul_rn(__dsub_rn(x2, ox), im_scale));
  o[4] = static_cast<float>(__dmul_rn(__dsub_rn(y2, oy), im_scale));
}

// wsl.py:139-155: labels_oh[0, g-1] = 1 for every ground-truth row (gt_classes > 0); labels_int32 = the LAST such
// row's class - 1 (the loop overwrites), -1 when the entry has none (the reference asserts; the host mirror raises).
__global__ void image_labels_kernel(const int32_t* __restrict__ gt, int n, int num_fg, float* __restrict__ labels_oh,
                                    int32_t* __restrict__ labels_int32) {
  __shared__ int last_row;
  if (threadIdx.x == 0) last_row = -1;
  for (int c = threadIdx.x; c < num_fg; c += blockDim.x) labels_oh[c] = 0.0f;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int g = gt[i];
    if (g > 0 && g <= num_fg) {
      labels_oh[g - 1] = 1.0f;                   // racing writers store the same value
      atomicMax(&last_row, i);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && labels_int32) labels_int32[0] = last_row >= 0 ? gt[last_row] - 1 : -1;
}

// loader_wsl.py:158-164: out = 0; out += lam0 * x0; out += lam1 * x1, every step rounded to float32 (no contraction:
// the build enables fmad).  Adding into the zero array turns a -0.0 product into +0.0, so the first add is kept.
__device__ __forceinline__ float mix1(float a, float b, float l0, float l1) {
  return __fadd_rn(__fadd_rn(0.0f, __fmul_rn(l0, a)), __fmul_rn(l1, b));
}
__global__ void bagging_mixup_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int64_t n, float l0, float l1,
                                     float* __restrict__ out, int vec_ok) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = vec_ok ? n / 4 : 0;
  for (int64_t i = tid; i < n4; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(x0)[i], b = reinterpret_cast<const float4*>(x1)[i];
    reinterpret_cast<float4*>(out)[i] = make_float4(mix1(a.x, b.x, l0, l1), mix1(a.y, b.y, l0, l1), mix1(a.z, b.z, l0, l1),
                                                    mix1(a.w, b.w, l0, l1));
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += stride) out[i] = mix1(x0[i], x1[i], l0, l1);
}

__global__ void set_column_kernel(float* __restrict__ a, int rows, int64_t ld, int col, float v) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) a[(size_t)r * ld + col] = v;
}

// convert_mcg.py:45-46: astype(np.uint16) - 1 (uint16 arithmetic: 0 wraps to 65535), columns (1, 0, 3, 2)
__global__ void convert_mcg_kernel(const double* __restrict__ in, int R, uint16_t* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  uint16_t v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = static_cast<uint16_t>(static_cast<uint16_t>(static_cast<long long>(in[4 * (size_t)r + k])) - 1u);
  *reinterpret_cast<uint2*>(out + 4 * (size_t)r) = make_uint2((uint32_t)v[1] | ((uint32_t)v[0] << 16), (uint32_t)v[3] | ((uint32_t)v[2] << 16));
}

}  // namespace
}  // namespace nawsod

using namespace nawsod;

extern "C" int nawsod_sample_rois(const float* boxes, int R, double im_scale, int crop_x1, int crop_y1, int crop_x2, int crop_y2,
                                  int batch_idx, float* rois, const float* obn_scores, float* obn_out, void* stream) {
  NAWSOD_REQUIRE(R >= 0, NAWSOD_ERR_SHAPE, "sample_rois: negative RoI count");
  if (R == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(boxes && rois, NAWSOD_ERR_ARG, "sample_rois: null boxes / rois");
  NAWSOD_REQUIRE(aligned16(boxes), NAWSOD_ERR_ALIGN, "sample_rois: boxes must be 16-byte aligned");
  NAWSOD_REQUIRE((obn_scores == nullptr) == (obn_out == nullptr), NAWSOD_ERR_ARG, "sample_rois: obn_scores and obn_out go together");
  NAWSOD_REQUIRE(crop_x2 >= crop_x1 && crop_y2 >= crop_y1, NAWSOD_ERR_ARG, "sample_rois: empty crop window (%d,%d,%d,%d)", crop_x1,
                 crop_y1, crop_x2, crop_y2);
  const int lim = 1 << 24;                        // the window is compared in float32 like the float32 box columns
  NAWSOD_REQUIRE(std::abs(crop_x1) < lim && std::abs(crop_y1) < lim && std::abs(crop_x2) < lim && std::abs(crop_y2) < lim, NAWSOD_ERR_ARG,
                 "sample_rois: crop coordinates beyond 2^24");
  sample_rois_kernel<<<(R + kThreads - 1) / kThreads, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      boxes, R, im_scale, (float)crop_x1, (float)crop_y1, (float)crop_x2, (float)crop_y2, (float)batch_idx, rois, obn_scores, obn_out);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_image_labels(const int32_t* gt_classes, int n, int num_classes, float* labels_oh, int32_t* labels_int32,
                                   void* stream) {
  NAWSOD_REQUIRE(n >= 0 && num_classes >= 2, NAWSOD_ERR_SHAPE, "image_labels: need n >= 0 rows and num_classes >= 2 (background + 1)");
  NAWSOD_REQUIRE(labels_oh && (gt_classes || n == 0), NAWSOD_ERR_ARG, "image_labels: null pointer");
  image_labels_kernel<<<1, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(gt_classes, n, num_classes - 1, labels_oh, labels_int32);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_bagging_mixup(const float* x0, const float* x1, int64_t n, float lam0, float lam1, float* out, void* stream) {
  NAWSOD_REQUIRE(n >= 0, NAWSOD_ERR_SHAPE, "bagging_mixup: negative size");
  if (n == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(x0 && x1 && out, NAWSOD_ERR_ARG, "bagging_mixup: null pointer");
  const int vec_ok = aligned16(x0) && aligned16(x1) && aligned16(out);
  const int64_t work = vec_ok ? (n + 3) / 4 : n;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((work + kThreads - 1) / kThreads, 8LL * sm_count()));
  bagging_mixup_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(x0, x1, n, lam0, lam1, out, vec_ok);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_set_column(float* a, int rows, int64_t ld, int col, float value, void* stream) {
  NAWSOD_REQUIRE(rows >= 0 && ld > 0 && col >= 0 && col < ld, NAWSOD_ERR_SHAPE, "set_column: column %d outside a row of %lld", col, (long long)ld);
  if (rows == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(a, NAWSOD_ERR_ARG, "set_column: null pointer");
  set_column_kernel<<<(rows + kThreads - 1) / kThreads, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, rows, ld, col, value);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_convert_mcg_boxes(const double* bboxes, int R, uint16_t* boxes_out, void* stream) {
  NAWSOD_REQUIRE(R >= 0, NAWSOD_ERR_SHAPE, "convert_mcg_boxes: negative box count");
  if (R == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(bboxes && boxes_out, NAWSOD_ERR_ARG, "convert_mcg_boxes: null pointer");
  NAWSOD_REQUIRE((reinterpret_cast<uintptr_t>(boxes_out) & 7u) == 0, NAWSOD_ERR_ALIGN, "convert_mcg_boxes: output must be 8-byte aligned");
  convert_mcg_kernel<<<(R + kThreads - 1) / kThreads, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(bboxes, R, boxes_out);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}
