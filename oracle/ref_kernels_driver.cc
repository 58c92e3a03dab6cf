// TEST INFRASTRUCTURE ONLY.  C ABI (for ctypes) around the reference's own CUDA kernels run on the host through
// oracle/cuda_host_shim.h.  The kernel text is NOT in this repository: oracle/build_ref_kernels.py cuts the
// anonymous-namespace block out of each reference file where it lies under /root/reference and writes it to
// oracle/_ref/gen/ (git-ignored), from where it is included below.
//   gen/roi_iou_kernels.inc          <- detectron/ops/roi_iou_op.cu            (iou<T>)
//   gen/roi_loop_pool_kernels.inc    <- detectron/ops/roi_loop_pool_op.cu      (ROIPoolForward<T>, ROIPoolBackward<T>)
//   gen/min_entropy_loss_kernels.inc <- detectron/ops/min_entropy_loss_op.cu   (get_norm_kernel, Forward, Backward)
#include <cstdint>
#include <cstring>
#include "cuda_host_shim.h"

namespace refk_iou {
using namespace cuda_host;
#include "gen/roi_iou_kernels.inc"
}  // namespace refk_iou

namespace refk_pool {
using namespace cuda_host;
#include "gen/roi_loop_pool_kernels.inc"
}  // namespace refk_pool

namespace refk_me {
using namespace cuda_host;
#include "gen/min_entropy_loss_kernels.inc"
}  // namespace refk_me

extern "C" {

// RoIIoUOp<float, CUDAContext>::RunOnDevice (roi_iou_op.cu:66-84): iou<float>(n * n, R, n, J)
void nawsod_refk_roi_iou(const float* rois, int n, float* J) { refk_iou::iou<float>(n * n, rois, n, J); }

// RoILoopPoolOp<float, CUDAContext>::RunOnDevice (roi_loop_pool_op.cu:145-187); rois9 is [R, 9]:
// (batch, outer x1 y1 x2 y2, inner x1 y1 x2 y2)
void nawsod_refk_roi_loop_pool_fwd(const float* X, const float* rois9, int R, int C, int H, int W, int PH, int PW,
                                   float spatial_scale, float* Y, int32_t* argmax) {
  refk_pool::ROIPoolForward<float>(R * C * PH * PW, X, spatial_scale, C, H, W, PH, PW, rois9, Y, argmax);
}

// RoILoopPoolGradientOp (roi_loop_pool_op.cu:189-224): zero-fill, then ROIPoolBackward
void nawsod_refk_roi_loop_pool_bwd(const float* dY, const int32_t* argmax, const float* rois9, int R, int N, int C, int H,
                                   int W, int PH, int PW, float spatial_scale, float* dX) {
  std::memset(dX, 0, sizeof(float) * (size_t)N * C * H * W);
  refk_pool::ROIPoolBackward<float>(R * C * PH * PW, dY, argmax, R, spatial_scale, C, H, W, PH, PW, dX, rois9);
}

// min_entropy_loss_op.cu:34-49: accumulates the loss terms into Y[0] and the selected-element count into norm[0]
void nawsod_refk_min_entropy_fwd(const float* X, const float* L, int N, int C, int B, float log_threshold, float* Y,
                                 float* norm) {
  refk_me::Forward<float>(N * C, X, L, N, C, B, log_threshold, Y, norm);
}
void nawsod_refk_min_entropy_norm(const float* X, const float* L, int N, int C, int B, float* norm) {
  refk_me::get_norm_kernel<float>(N * C, X, L, N, C, B, norm);
}
// min_entropy_loss_op.cu:52-66: dX only at the selected classes (the op zero-fills dX first, :133-135)
void nawsod_refk_min_entropy_bwd(const float* X, const float* L, int N, int C, int B, const float* scale,
                                 float log_threshold, float diff_threshold, float* dX) {
  std::memset(dX, 0, sizeof(float) * (size_t)N * C);
  refk_me::Backward<float>(N * C, X, L, N, C, B, scale, log_threshold, diff_threshold, dX);
}

}  // extern "C"
