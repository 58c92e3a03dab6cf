/* CPU oracle (TEST INFRASTRUCTURE ONLY): plain-C restatement of the reference's greedy NMS,
 * detectron/utils/cython_nms.pyx:38-93.  The visiting order is an INPUT (the caller passes
 * scores.argsort()[::-1] exactly as the reference computes it, cython_nms.pyx:45), all box
 * arithmetic is float32 operation by operation (build with -ffp-contract=off), and a box is
 * suppressed when ovr >= thresh (:86).  Returns the number of kept boxes; keep[i] = 1 for the
 * survivors in ORIGINAL row order (np.where(suppressed == 0)[0], :93). */
#include <stdint.h>

static float fmax_ref(float a, float b) { return a >= b ? a : b; }   /* cython_nms.pyx:28-29 */
static float fmin_ref(float a, float b) { return a <= b ? a : b; }   /* cython_nms.pyx:31-32 */

int nawsod_oracle_nms(const float* dets /* [n,5] x1 y1 x2 y2 score */, int n, const int64_t* order,
                      float thresh, uint8_t* keep) {
  int kept = 0;
  for (int i = 0; i < n; ++i) keep[i] = 1;
  for (int _i = 0; _i < n; ++_i) {
    const int i = (int)order[_i];
    if (!keep[i]) continue;
    ++kept;
    const float ix1 = dets[i * 5 + 0], iy1 = dets[i * 5 + 1], ix2 = dets[i * 5 + 2], iy2 = dets[i * 5 + 3];
    const float iarea = (ix2 - ix1 + 1) * (iy2 - iy1 + 1);                 /* :44 */
    for (int _j = _i + 1; _j < n; ++_j) {
      const int j = (int)order[_j];
      if (!keep[j]) continue;
      const float jx1 = dets[j * 5 + 0], jy1 = dets[j * 5 + 1], jx2 = dets[j * 5 + 2], jy2 = dets[j * 5 + 3];
      const float jarea = (jx2 - jx1 + 1) * (jy2 - jy1 + 1);
      const float xx1 = fmax_ref(ix1, jx1), yy1 = fmax_ref(iy1, jy1);     /* :77-80 */
      const float xx2 = fmin_ref(ix2, jx2), yy2 = fmin_ref(iy2, jy2);
      const float w = fmax_ref(0.0f, xx2 - xx1 + 1), h = fmax_ref(0.0f, yy2 - yy1 + 1);   /* :81-82 */
      const float inter = w * h;                                           /* :83 */
      const float ovr = inter / (iarea + jarea - inter);                   /* :84 */
      if (ovr >= thresh) keep[j] = 0;                                      /* :85-86 */
    }
  }
  return kept;
}
