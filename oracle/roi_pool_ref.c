/* Plain-C restatement of Caffe2 RoIPoolF / RoIPoolFGradient (TEST INFRASTRUCTURE ONLY).
 *
 * RoIPoolF's source (pytorch v1.3.0 modules/detectron/roi_pool_f_op.cu) is not in the
 * reference tree; the arithmetic follows the in-tree clone
 * /root/reference/detectron/ops/roi_loop_pool_op.cu:19-140 with RoIPoolF's three deltas
 * (SURVEY.md row a1): rois stride 5 (:38), no inner rectangle (:80-85), and
 * maxval = is_empty ? 0 : -FLT_MAX (the commented original, :72).  One loop iteration here
 * is one CUDA thread there.  Used by tests/ as the checker at full sizes and by bench.py's
 * cpu_baseline leg; never linked into the product library.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* X [N,C,H,W], rois [R,5], Y/argmax [R,C,PH,PW]; argmax may be NULL (is_test). */
void nawsod_oracle_roi_pool_f(const float* X, const float* rois, int N, int C, int H, int W, int R,
                              float spatial_scale, int PH, int PW, float* Y, int32_t* argmax) {
  (void)N;
#pragma omp parallel for schedule(dynamic, 4)
  for (int n = 0; n < R; ++n) {
    const float* roi = rois + (size_t)n * 5;                       /* :38 (stride 5) */
    int roi_batch_ind = (int)roi[0];                               /* :39 */
    int roi_start_w = (int)roundf(roi[1] * spatial_scale);         /* :42-45 */
    int roi_start_h = (int)roundf(roi[2] * spatial_scale);
    int roi_end_w = (int)roundf(roi[3] * spatial_scale);
    int roi_end_h = (int)roundf(roi[4] * spatial_scale);
    int roi_width = imax(roi_end_w - roi_start_w + 1, 1);          /* :54-55 */
    int roi_height = imax(roi_end_h - roi_start_h + 1, 1);
    float bin_size_h = (float)roi_height / (float)PH;              /* :56-57 */
    float bin_size_w = (float)roi_width / (float)PW;
    for (int c = 0; c < C; ++c) {
      const float* plane = X + ((size_t)roi_batch_ind * C + c) * H * W;   /* :76-77 */
      for (int ph = 0; ph < PH; ++ph) {
        for (int pw = 0; pw < PW; ++pw) {
          int hstart = (int)floorf((float)ph * bin_size_h);        /* :59-62 */
          int wstart = (int)floorf((float)pw * bin_size_w);
          int hend = (int)ceilf((float)(ph + 1) * bin_size_h);
          int wend = (int)ceilf((float)(pw + 1) * bin_size_w);
          hstart = imin(imax(hstart + roi_start_h, 0), H);         /* :65-68 */
          hend = imin(imax(hend + roi_start_h, 0), H);
          wstart = imin(imax(wstart + roi_start_w, 0), W);
          wend = imin(imax(wend + roi_start_w, 0), W);
          int is_empty = (hend <= hstart) || (wend <= wstart);     /* :69 */
          float maxval = is_empty ? 0.f : -FLT_MAX;                /* :72 (RoIPoolF) */
          int maxidx = -1;                                         /* :76 */
          for (int h = hstart; h < hend; ++h) {
            for (int w = wstart; w < wend; ++w) {
              int bottom_index = h * W + w;                        /* :88 */
              if (plane[bottom_index] > maxval) {                  /* :89 strict > */
                maxval = plane[bottom_index];
                maxidx = bottom_index;
              }
            }
          }
          size_t index = (((size_t)n * C + c) * PH + ph) * PW + pw;
          Y[index] = maxval;                                       /* :97 */
          if (argmax) argmax[index] = maxidx;                      /* :98-100 */
        }
      }
    }
  }
}

/* dX [N,C,H,W] is zero-filled here (:199-201), then dX[b,c,argmax] += dY (:134-139).
 * Sequential (r, c, ph, pw) order: the reference's atomicAdd order is unspecified. */
void nawsod_oracle_roi_pool_f_grad(const float* dY, const int32_t* argmax, const float* rois, int N, int C,
                                   int H, int W, int R, int PH, int PW, float* dX) {
  memset(dX, 0, (size_t)N * C * H * W * sizeof(float));
#pragma omp parallel for schedule(static)
  for (int c = 0; c < C; ++c) {
    for (int n = 0; n < R; ++n) {
      int roi_batch_ind = (int)rois[(size_t)n * 5];
      float* plane = dX + ((size_t)roi_batch_ind * C + c) * H * W;
      size_t top = ((size_t)n * C + c) * PH * PW;
      for (int k = 0; k < PH * PW; ++k) {
        int a = argmax[top + k];
        if (a != -1) plane[a] += dY[top + k];
      }
    }
  }
}
