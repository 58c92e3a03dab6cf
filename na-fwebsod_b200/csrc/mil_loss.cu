// Fused two-stream MIL head + noise-aware class weights + weighted multi-label CE, forward and
// backward to the fc8 logits, as ONE cooperative kernel (rows a5..a9 of SURVEY.md section 8).
//
// Replaces ~60 Caffe2 operators per step: Softmax/Transpose/Softmax/Transpose/Mul
// (detectron/modeling/wsl_heads.py:49-55), the noise stream Add + softmaxes
// (detectron/modeling/webly_heads.py:57-74), ReduceSum (wsl_heads.py:227), RoIIoU + the
// spatial-entropy weight graph (webly_heads.py:265-391, detectron/ops/roi_iou_op.cu:28-62), the two
// WeightedCrossEntropyWithLogits ops that the reference runs on the CPU through GPUFallbackOp
// (detectron/ops/cross_entropy_wsl_op.cc:88-180, .cu:365-369) and all their gradient ops.
//
// Structure: a persistent cooperative grid (<= one CTA per SM).  Every image owns a contiguous
// group of CTAs and every CTA a contiguous block of that image's RoIs.  Phases, separated by
// grid barriers:
//   P1  per-CTA partial (max, sum-exp) of the RoI-axis softmax columns (both streams)
//   P2  combine partials; per-row class softmax (warp-shuffle reductions), P = a_cls*a_det,
//       E = -P log P, per-CTA partial column sums y
//   P3  D = J.E without materialising the R x R IoU matrix: lanes = rows i, warps = j-slices,
//       E[j,:] and box j arrive by broadcast loads; E^2/D column partial sums
//   P4  class weights, losses (sequential double accumulation like the reference), dL/dy
//   P5  analytic backward: d_fc8c = P*dy - a_cls*sum_c(dy*P),  d_fc8d = a_det*(dy*a_cls - dy*y)
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace nawsod {
namespace {

constexpr int kMaxC = 128;
constexpr int kMaxB = 64;
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kQ = kMaxC / 32;   // class slots per lane
constexpr int kMaxGrid = 1024;

struct MilParams {
  const float *fc8c, *fc8d, *nfc8c, *nfc8d, *rois, *labels;
  const int32_t* roi_offsets;
  int R, C, B, flags;
  long long ldl, ldg;   // row pitch of the logit inputs / logit-gradient outputs
  float *rois_pred, *cls_prob, *rois_pred_noise, *cls_prob_noise, *class_weight, *class_weight_noise, *loss;
  float *d_fc8c, *d_fc8d, *d_nfc8c, *d_nfc8d;
  // workspace
  float* part_col;   // [2][G][C][2]
  float* part_y;     // [2][G][C]
  float* part_hat;   // [G][C]
  float* E;          // [R][CP]
  int4* box;         // [R]
  int CP;
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void combine_ms(float& m, float& s, float m2, float s2) {
  const float mm = fmaxf(m, m2);
  if (mm == -INFINITY) { m = mm; s = 0.f; return; }
  s = s * expf(m - mm) + s2 * expf(m2 - mm);
  m = mm;
}

// One row of one stream: class softmax over C (warp reduction) and the RoI-axis softmax value
// from the column statistics.  Lane owns classes lane, lane+32, ...
struct RowProbs { float a_cls[kQ], a_det[kQ], P[kQ]; };

__device__ __forceinline__ void row_probs(const float* __restrict__ lc, const float* __restrict__ nlc,
                                          const float* __restrict__ ld, const float* __restrict__ nld, int r, int C, long long pitch,
                                          const float* colmax, const float* colsum, int lane, RowProbs& o) {
  float xc[kQ], xd[kQ];
  float m = -INFINITY;
#pragma unroll
  for (int q = 0; q < kQ; ++q) {
    const int c = lane + 32 * q;
    if (c < C) {
      xc[q] = lc[(size_t)r * pitch + c];
      xd[q] = ld[(size_t)r * pitch + c];
      if (nlc) { xc[q] += nlc[(size_t)r * pitch + c]; xd[q] += nld[(size_t)r * pitch + c]; }
      m = fmaxf(m, xc[q]);
    } else { xc[q] = -INFINITY; xd[q] = -INFINITY; }
  }
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < kQ; ++q) {
    const int c = lane + 32 * q;
    o.a_cls[q] = (c < C) ? expf(xc[q] - m) : 0.f;
    s += o.a_cls[q];
  }
  s = warp_sum(s);
#pragma unroll
  for (int q = 0; q < kQ; ++q) {
    const int c = lane + 32 * q;
    if (c < C) {
      o.a_cls[q] = o.a_cls[q] / s;
      o.a_det[q] = expf(xd[q] - colmax[c]) / colsum[c];
      o.P[q] = o.a_cls[q] * o.a_det[q];
    } else { o.a_cls[q] = 0.f; o.a_det[q] = 0.f; o.P[q] = 0.f; }
  }
}

// D rows for up to 32 RoIs (lanes) against the j-slice of this warp.  CT = class tile held in registers.
template <int CT>
__device__ __forceinline__ void je_tile(const MilParams& p, int img_row0, int img_rows, int i_row, bool i_valid,
                                        int c0, int warp, float* D_s /* [32][kMaxC+1] */, int lane) {
  float acc[CT];
#pragma unroll
  for (int k = 0; k < CT; ++k) acc[k] = 0.f;
  int4 bi = make_int4(0, 0, 0, 0);
  double area_i = 0.0;
  if (i_valid) {
    bi = p.box[i_row];
    area_i = (double)(bi.z - bi.x + 1) * (double)(bi.w - bi.y + 1);
  }
  const int per = (img_rows + kWarps - 1) / kWarps;
  const int j0 = img_row0 + warp * per;
  const int j1 = min(img_row0 + img_rows, j0 + per);
  for (int j = j0; j < j1; ++j) {
    const int4 bj = p.box[j];                                    // broadcast load
    // detectron/ops/roi_iou_op.cu:37-60
    const int xmin = max(bi.x, bj.x), ymin = max(bi.y, bj.y);
    const int xmax = min(bi.z, bj.z), ymax = min(bi.w, bj.w);
    const int w = max(xmax - xmin + 1, 0), h = max(ymax - ymin + 1, 0);
    const float inters = static_cast<float>(w * h);
    const double area_j = (double)(bj.z - bj.x + 1) * (double)(bj.w - bj.y + 1);
    const float uni = static_cast<float>(area_i + area_j - (double)inters);
    float J = __fdiv_rn(inters, uni);
    if (j == i_row) J = 1.0f;
    const float4* Ej = reinterpret_cast<const float4*>(p.E + (size_t)j * p.CP + c0);
#pragma unroll
    for (int k = 0; k < CT / 4; ++k) {
      const float4 e = Ej[k];                                    // broadcast load
      acc[4 * k + 0] = fmaf(J, e.x, acc[4 * k + 0]);
      acc[4 * k + 1] = fmaf(J, e.y, acc[4 * k + 1]);
      acc[4 * k + 2] = fmaf(J, e.z, acc[4 * k + 2]);
      acc[4 * k + 3] = fmaf(J, e.w, acc[4 * k + 3]);
    }
  }
  if (i_valid) {
#pragma unroll
    for (int k = 0; k < CT; ++k)
      if (c0 + k < p.C) atomicAdd(&D_s[lane * (kMaxC + 1) + c0 + k], acc[k]);
  }
}

__global__ void __launch_bounds__(kThreads, 1) mil_head_kernel(const MilParams p) {
  cg::grid_group grid = cg::this_grid();
  const int G = gridDim.x, g = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, B = p.B, R = p.R;
  const bool noise = p.nfc8c != nullptr;
  const bool entropy = noise && (p.flags & NAWSOD_MIL_ENTROPY);
  const int nstreams = noise ? 2 : 1;

  __shared__ float colmax_s[2][kMaxC], colsum_s[2][kMaxC], y_s[2][kMaxC], dy_s[2][kMaxC], hat_s[kMaxC];
  __shared__ float w_s[2][kMaxC];
  __shared__ float red_m[kThreads], red_s[kThreads];
  __shared__ double term_s[2][kMaxC];
  __shared__ int first_cta_s[kMaxB + 1];
  extern __shared__ float D_s[];   // [32][kMaxC+1], only when entropy

  // ---- CTA -> (image, row block) -------------------------------------------------------------
  if (tid <= B) {
    const long long off = p.roi_offsets[tid];
    first_cta_s[tid] = tid + (int)(((long long)(G - B) * off) / max(R, 1));
  }
  __syncthreads();
  int b = 0;
  while (b + 1 < B && g >= first_cta_s[b + 1]) ++b;
  const int cta0 = first_cta_s[b], Gb = first_cta_s[b + 1] - cta0, k = g - cta0;
  const int img_row0 = p.roi_offsets[b], img_rows = p.roi_offsets[b + 1] - img_row0;
  // CTAs past the last image's group own no rows (roi_offsets[B] may be smaller than R when the caller
  // pads the row count, e.g. de-duplicated test-time RoIs); they only take part in the grid barriers
  const bool idle_cta = k >= Gb;
  const int row0 = idle_cta ? img_row0 + img_rows : img_row0 + (int)(((long long)img_rows * k) / Gb);
  const int row1 = idle_cta ? img_row0 + img_rows : img_row0 + (int)(((long long)img_rows * (k + 1)) / Gb);

  // ---- P1: partial column statistics of the RoI-axis softmax -----------------------------------
  {
    const int KR = kThreads / C;
    for (int s = 0; s < nstreams; ++s) {
      const float* ld = p.fc8d;
      const float* nld = s ? p.nfc8d : nullptr;
      float m = -INFINITY, sum = 0.f;
      const int col = tid % C, rl = tid / C;
      if (rl < KR) {
        for (int r = row0 + rl; r < row1; r += KR) {
          float v = ld[(size_t)r * p.ldl + col];
          if (nld) v += nld[(size_t)r * p.ldl + col];
          if (v > m) { sum = sum * expf(m - v) + 1.f; m = v; }
          else sum += expf(v - m);
        }
      }
      red_m[tid] = m; red_s[tid] = sum;
      __syncthreads();
      if (tid < C) {
        float mm = -INFINITY, ss = 0.f;
        for (int q = 0; q < KR; ++q) combine_ms(mm, ss, red_m[q * C + tid], red_s[q * C + tid]);
        float* dst = p.part_col + (((size_t)s * G + g) * C + tid) * 2;
        dst[0] = mm; dst[1] = ss;
      }
      __syncthreads();
    }
  }
  grid.sync();

  // ---- P2: probabilities, E, partial y ---------------------------------------------------------
  for (int i = tid; i < nstreams * C; i += kThreads) {
    const int s = i / C, c = i - s * C;
    float mm = -INFINITY, ss = 0.f;
    for (int q = 0; q < Gb; ++q) {
      const float* src = p.part_col + (((size_t)s * G + cta0 + q) * C + c) * 2;
      combine_ms(mm, ss, src[0], src[1]);
    }
    colmax_s[s][c] = mm; colsum_s[s][c] = ss;
    y_s[s][c] = 0.f;
  }
  for (int i = tid; i < C; i += kThreads) hat_s[i] = 0.f;
  __syncthreads();
  {
    float yacc[2][kQ];
#pragma unroll
    for (int q = 0; q < kQ; ++q) { yacc[0][q] = 0.f; yacc[1][q] = 0.f; }
    for (int r = row0 + warp; r < row1; r += kWarps) {
      RowProbs pr;
      row_probs(p.fc8c, nullptr, p.fc8d, nullptr, r, C, p.ldl, colmax_s[0], colsum_s[0], lane, pr);
#pragma unroll
      for (int q = 0; q < kQ; ++q) {
        const int c = lane + 32 * q;
        if (c < C) {
          yacc[0][q] += pr.P[q];
          if (p.rois_pred) p.rois_pred[(size_t)r * C + c] = pr.P[q];
          if (entropy) {
            // webly_heads.py:276-279: E = -(P * log P), NaN -> 0
            float e = -(pr.P[q] * logf(pr.P[q]));
            if (e != e) e = 0.f;
            p.E[(size_t)r * p.CP + c] = e;
          }
        } else if (entropy && c < p.CP) {
          p.E[(size_t)r * p.CP + c] = 0.f;
        }
      }
      if (entropy && lane == 0) {
        const float* roi = p.rois + (size_t)r * 5;
        p.box[r] = make_int4((int)roi[1], (int)roi[2], (int)roi[3], (int)roi[4]);   // float -> int truncation
      }
      if (noise) {
        row_probs(p.fc8c, p.nfc8c, p.fc8d, p.nfc8d, r, C, p.ldl, colmax_s[1], colsum_s[1], lane, pr);
#pragma unroll
        for (int q = 0; q < kQ; ++q) {
          const int c = lane + 32 * q;
          if (c < C) {
            yacc[1][q] += pr.P[q];
            if (p.rois_pred_noise) p.rois_pred_noise[(size_t)r * C + c] = pr.P[q];
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      const int c = lane + 32 * q;
      if (c < C) {
        atomicAdd(&y_s[0][c], yacc[0][q]);
        if (noise) atomicAdd(&y_s[1][c], yacc[1][q]);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < nstreams * C; i += kThreads) {
    const int s = i / C, c = i - s * C;
    p.part_y[((size_t)s * G + g) * C + c] = y_s[s][c];
  }
  grid.sync();

  // ---- P3: D = J.E, hatE = E*E/D, partial column sums --------------------------------------------
  if (entropy) {
    for (int base = row0; base < row1; base += 32) {
      for (int i = tid; i < 32 * (kMaxC + 1); i += kThreads) D_s[i] = 0.f;
      __syncthreads();
      const int i_row = base + lane;
      const bool i_valid = i_row < row1;
      for (int c0 = 0; c0 < C; c0 += 80) {
        if (C - c0 <= 24) je_tile<24>(p, img_row0, img_rows, i_row, i_valid, c0, warp, D_s, lane);
        else if (C - c0 <= 48) je_tile<48>(p, img_row0, img_rows, i_row, i_valid, c0, warp, D_s, lane);
        else je_tile<80>(p, img_row0, img_rows, i_row, i_valid, c0, warp, D_s, lane);
      }
      __syncthreads();
      const int nrows = min(32, row1 - base);
      for (int i = tid; i < nrows * C; i += kThreads) {
        const int rl = i / C, c = i - rl * C;
        float D = D_s[rl * (kMaxC + 1) + c];
        D = D >= 0.f ? D : 0.01f * D;                              // LeakyRelu, webly_heads.py:281
        const float E = p.E[(size_t)(base + rl) * p.CP + c];
        const float Gq = __fdiv_rn(E, D);                           // :282
        atomicAdd(&hat_s[c], E * Gq);                               // :283, :285-288
      }
      __syncthreads();
    }
    for (int i = tid; i < C; i += kThreads) p.part_hat[(size_t)g * C + i] = hat_s[i];
    grid.sync();
  }

  // ---- P4: image scores, class weights, losses, dL/dy ------------------------------------------
  const bool is_mean = (p.flags & NAWSOD_MIL_MEAN) != 0;
  const float norm = is_mean ? static_cast<float>(C) : 1.f;
  for (int i = tid; i < nstreams * C; i += kThreads) {
    const int s = i / C, c = i - s * C;
    float acc = 0.f;
    for (int q = 0; q < Gb; ++q) acc += p.part_y[((size_t)s * G + cta0 + q) * C + c];
    y_s[s][c] = acc;
  }
  __syncthreads();
  for (int c = tid; c < C; c += kThreads) {
    const float L = p.labels[(size_t)b * C + c];
    float wn = 0.f, wc = 1.f;
    if (entropy) {
      float hs = 0.f;
      for (int q = 0; q < Gb; ++q) hs += p.part_hat[(size_t)(cta0 + q) * C + c];
      const float y = y_s[0][c];
      const float logy = logf(y);                                   // webly_heads.py:335
      const float logN = logf(static_cast<float>(img_rows));        // :336
      const float denom = (logN - logy) * y;                        // :337-340
      float nrm = __fdiv_rn(hs, denom);                             // :345-347
      nrm = fminf(fmaxf(nrm, 0.f), 1.f);                            // :350-353
      wn = nrm * (1.f - L);                                         // :368-371
      wc = 1.f - wn;                                                // :373-374
    }
    w_s[0][c] = wc; w_s[1][c] = wn;
    for (int s = 0; s < nstreams; ++s) {
      const float x = y_s[s][c];
      const float wgt = w_s[s][c];
      const bool weighted = entropy;
      // detectron/ops/cross_entropy_wsl_op.cc:122-127 (forward term, evaluated in double)
      const float prob = fmaxf(x, 1e-20f), one_prob = fmaxf(1.f - x, 1e-20f);
      double term = (double)L * log((double)prob) + (double)(1.f - L) * log((double)one_prob);
      if (weighted) term *= (double)wgt;
      term_s[s][c] = term;
      // :166-177 (gradient; seed dY = 1, N = 1): upper clamp before the weight
      float d = fminf((-1.f * L / prob - (-1.f) * (1.f - L) / one_prob) / norm, 1e4f);
      if (weighted) d *= wgt;
      dy_s[s][c] = d;
    }
  }
  __syncthreads();
  if (k == 0) {   // first CTA of the image publishes the per-image outputs
    for (int c = tid; c < C; c += kThreads) {
      if (p.cls_prob) p.cls_prob[(size_t)b * C + c] = y_s[0][c];
      if (noise && p.cls_prob_noise) p.cls_prob_noise[(size_t)b * C + c] = y_s[1][c];
      if (p.class_weight) p.class_weight[(size_t)b * C + c] = w_s[0][c];
      if (p.class_weight_noise) p.class_weight_noise[(size_t)b * C + c] = w_s[1][c];
    }
    if (tid < nstreams && p.loss) {
      float loss = 0.f;
      for (int c = 0; c < C; ++c) loss = static_cast<float>((double)loss - term_s[tid][c]);   // loss -= term
      p.loss[(size_t)b * 2 + tid] = loss / norm;
    }
    if (tid == 1 && !noise && p.loss) p.loss[(size_t)b * 2 + 1] = 0.f;
  }

  // ---- P5: backward to the logits -------------------------------------------------------------
  if (p.flags & NAWSOD_MIL_BACKWARD) {
    for (int r = row0 + warp; r < row1; r += kWarps) {
      RowProbs pr;
      row_probs(p.fc8c, nullptr, p.fc8d, nullptr, r, C, p.ldl, colmax_s[0], colsum_s[0], lane, pr);
      float S = 0.f;
#pragma unroll
      for (int q = 0; q < kQ; ++q) {
        const int c = lane + 32 * q;
        if (c < C) S += dy_s[0][c] * pr.P[q];
      }
      S = warp_sum(S);
      float gc[kQ], gd[kQ];
#pragma unroll
      for (int q = 0; q < kQ; ++q) {
        const int c = lane + 32 * q;
        if (c < C) {
          const float dy = dy_s[0][c];
          gc[q] = pr.a_cls[q] * (dy * pr.a_det[q] - S);
          gd[q] = pr.a_det[q] * (dy * pr.a_cls[q] - dy * y_s[0][c]);
        }
      }
      if (noise) {
        row_probs(p.fc8c, p.nfc8c, p.fc8d, p.nfc8d, r, C, p.ldl, colmax_s[1], colsum_s[1], lane, pr);
        float Sn = 0.f;
#pragma unroll
        for (int q = 0; q < kQ; ++q) {
          const int c = lane + 32 * q;
          if (c < C) Sn += dy_s[1][c] * pr.P[q];
        }
        Sn = warp_sum(Sn);
#pragma unroll
        for (int q = 0; q < kQ; ++q) {
          const int c = lane + 32 * q;
          if (c < C) {
            const float dy = dy_s[1][c];
            const float nc = pr.a_cls[q] * (dy * pr.a_det[q] - Sn);
            const float nd = pr.a_det[q] * (dy * pr.a_cls[q] - dy * y_s[1][c]);
            p.d_nfc8c[(size_t)r * p.ldg + c] = nc;
            p.d_nfc8d[(size_t)r * p.ldg + c] = nd;
            gc[q] += nc;      // Add fans the gradient out to both summands
            gd[q] += nd;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kQ; ++q) {
        const int c = lane + 32 * q;
        if (c < C) {
          p.d_fc8c[(size_t)r * p.ldg + c] = gc[q];
          p.d_fc8d[(size_t)r * p.ldg + c] = gd[q];
        }
      }
    }
  }
}

// ---- stand-alone operators ---------------------------------------------------------------------
__global__ void roi_iou_kernel(const float* __restrict__ rois, int n, float* __restrict__ J) {
  const size_t total = (size_t)n * n;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = static_cast<int>(idx % n), j = static_cast<int>(idx / n);
    if (i == j) { J[idx] = 1.0f; continue; }
    const int ixmin = rois[i * 5 + 1], iymin = rois[i * 5 + 2], ixmax = rois[i * 5 + 3], iymax = rois[i * 5 + 4];
    const int jxmin = rois[j * 5 + 1], jymin = rois[j * 5 + 2], jxmax = rois[j * 5 + 3], jymax = rois[j * 5 + 4];
    const int xmin = max(ixmin, jxmin), ymin = max(iymin, jymin);
    const int xmax = min(ixmax, jxmax), ymax = min(iymax, jymax);
    const int w = max(xmax - xmin + 1, 0), h = max(ymax - ymin + 1, 0);
    const float inters = static_cast<float>(w * h);
    const double ai = (double)(ixmax - ixmin + 1) * (double)(iymax - iymin + 1);
    const double aj = (double)(jxmax - jxmin + 1) * (double)(jymax - jymin + 1);
    const float uni = static_cast<float>(ai + aj - (double)inters);
    J[idx] = __fdiv_rn(inters, uni);
  }
}

// One block; thread c evaluates class terms, thread 0 accumulates in the reference's order.
__global__ void ce_fwd_kernel(const float* X, const float* L, const float* Wt, int N, int C, int is_mean, float* Y) {
  extern __shared__ double terms[];
  const int total = N * C;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const float prob = fmaxf(X[i], 1e-20f), one_prob = fmaxf(1.f - X[i], 1e-20f);
    double t = (double)L[i] * log((double)prob) + (double)(1.f - L[i]) * log((double)one_prob);
    if (Wt) t *= (double)Wt[i];
    terms[i] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float loss = 0.f;
    for (int i = 0; i < total; ++i) loss = static_cast<float>((double)loss - terms[i]);
    const float norm = is_mean ? static_cast<float>(C) : 1.f;
    Y[0] = (loss / norm) * static_cast<float>(1.0 / N);
  }
}

__global__ void ce_bwd_kernel(const float* X, const float* L, const float* Wt, const float* dY, int N, int C,
                              int is_mean, float* dX) {
  const int total = N * C;
  const float norm = is_mean ? static_cast<float>(C) : 1.f;
  const float scale = static_cast<float>(1.0 / N);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const float grad = dY[0];
    const float prob = fmaxf(X[i], 1e-20f), one_prob = fmaxf(1.f - X[i], 1e-20f);
    float d = fminf(grad * (-1.f * L[i] / prob - (-1.f) * (1.f - L[i]) / one_prob) / norm, 1e4f);
    if (Wt) d *= Wt[i];
    dX[i] = d * scale;
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct WsLayout { size_t part_col, part_y, part_hat, E, box, total; int CP; };
WsLayout ws_layout(int R, int C) {
  WsLayout w;
  // row pitch of E: whole register tiles of je_tile (24 / 48 / 80 classes) must stay inside a row
  w.CP = 0;
  for (int c0 = 0; c0 < C; c0 += 80) {
    const int rem = C - c0;
    w.CP = c0 + (rem <= 24 ? 24 : rem <= 48 ? 48 : 80);
  }
  size_t o = 0;
  w.part_col = o; o = align_up(o + (size_t)2 * kMaxGrid * C * 2 * sizeof(float), 256);
  w.part_y = o;   o = align_up(o + (size_t)2 * kMaxGrid * C * sizeof(float), 256);
  w.part_hat = o; o = align_up(o + (size_t)kMaxGrid * C * sizeof(float), 256);
  w.E = o;        o = align_up(o + (size_t)std::max(R, 1) * w.CP * sizeof(float), 256);
  w.box = o;      o = align_up(o + (size_t)std::max(R, 1) * sizeof(int4), 256);
  w.total = o;
  return w;
}

}  // namespace
}  // namespace nawsod

using namespace nawsod;

extern "C" int64_t nawsod_mil_workspace_bytes(int R, int C, int B) {
  (void)B;
  if (R < 0 || C <= 0) return 0;
  return (int64_t)ws_layout(R, C).total;
}

extern "C" int nawsod_mil_head_fwd_bwd(const float* fc8c, const float* fc8d, const float* nfc8c, const float* nfc8d,
                                       int64_t ld_logits, const float* rois, const int32_t* roi_offsets,
                                       const float* labels_oh, int R, int C, int B, int flags, float* rois_pred, float* cls_prob,
                                       float* rois_pred_noise, float* cls_prob_noise, float* class_weight,
                                       float* class_weight_noise, float* loss, float* d_fc8c, float* d_fc8d,
                                       float* d_nfc8c, float* d_nfc8d, int64_t ld_grads, void* workspace, void* stream) {
  NAWSOD_REQUIRE(R > 0 && C > 0 && B > 0, NAWSOD_ERR_SHAPE, "mil_head: need R, C, B > 0 (got %d, %d, %d)", R, C, B);
  NAWSOD_REQUIRE(C <= kMaxC, NAWSOD_ERR_UNSUPPORTED, "mil_head: C=%d > %d", C, kMaxC);
  NAWSOD_REQUIRE(B <= kMaxB, NAWSOD_ERR_UNSUPPORTED, "mil_head: B=%d > %d", B, kMaxB);
  NAWSOD_REQUIRE(fc8c && fc8d && rois && roi_offsets && labels_oh && workspace, NAWSOD_ERR_ARG,
                 "mil_head: null input pointer");
  NAWSOD_REQUIRE((nfc8c == nullptr) == (nfc8d == nullptr), NAWSOD_ERR_ARG,
                 "mil_head: nfc8c and nfc8d must both be given or both be NULL");
  if (flags & NAWSOD_MIL_BACKWARD) {
    NAWSOD_REQUIRE(d_fc8c && d_fc8d && (!nfc8c || (d_nfc8c && d_nfc8d)), NAWSOD_ERR_ARG,
                   "mil_head: BACKWARD needs the d_* outputs");
  }
  NAWSOD_REQUIRE(aligned16(workspace), NAWSOD_ERR_ALIGN, "mil_head: workspace must be 16-byte aligned");
  NAWSOD_REQUIRE(ld_logits >= C && (!(flags & NAWSOD_MIL_BACKWARD) || ld_grads >= C), NAWSOD_ERR_SHAPE,
                 "mil_head: leading dimensions must be >= C");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const WsLayout w = ws_layout(R, C);
  char* base = static_cast<char*>(workspace);
  MilParams p;
  p.fc8c = fc8c; p.fc8d = fc8d; p.nfc8c = nfc8c; p.nfc8d = nfc8d; p.rois = rois; p.labels = labels_oh;
  p.roi_offsets = roi_offsets; p.R = R; p.C = C; p.B = B; p.flags = flags; p.ldl = ld_logits; p.ldg = ld_grads;
  p.rois_pred = rois_pred; p.cls_prob = cls_prob; p.rois_pred_noise = rois_pred_noise;
  p.cls_prob_noise = cls_prob_noise; p.class_weight = class_weight; p.class_weight_noise = class_weight_noise;
  p.loss = loss; p.d_fc8c = d_fc8c; p.d_fc8d = d_fc8d; p.d_nfc8c = d_nfc8c; p.d_nfc8d = d_nfc8d;
  p.part_col = reinterpret_cast<float*>(base + w.part_col);
  p.part_y = reinterpret_cast<float*>(base + w.part_y);
  p.part_hat = reinterpret_cast<float*>(base + w.part_hat);
  p.E = reinterpret_cast<float*>(base + w.E);
  p.box = reinterpret_cast<int4*>(base + w.box);
  p.CP = w.CP;

  const bool entropy = nfc8c && (flags & NAWSOD_MIL_ENTROPY);
  const size_t dyn_smem = entropy ? (size_t)32 * (kMaxC + 1) * sizeof(float) : 0;
  int grid = (int)get_tuning("mil_ctas", 0);
  if (grid <= 0) grid = sm_count();
  grid = std::max(B, std::min(grid, std::min(kMaxGrid, std::max(B, (R + 7) / 8))));
  int max_blocks = 0;
  NAWSOD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, mil_head_kernel, kThreads, dyn_smem));
  NAWSOD_REQUIRE(max_blocks >= 1, NAWSOD_ERR_CUDA, "mil_head: kernel does not fit on an SM");
  grid = std::min(grid, max_blocks * sm_count());
  NAWSOD_REQUIRE(grid >= B, NAWSOD_ERR_UNSUPPORTED, "mil_head: more images (%d) than co-resident CTAs (%d)", B, grid);
  void* args[] = {(void*)&p};
  NAWSOD_CUDA_OK(cudaLaunchCooperativeKernel((const void*)mil_head_kernel, dim3(grid), dim3(kThreads), args, dyn_smem, st));
  return NAWSOD_OK;
}

extern "C" int nawsod_roi_iou(const float* rois, int R, float* J, void* stream) {
  NAWSOD_REQUIRE(R >= 0, NAWSOD_ERR_SHAPE, "roi_iou: negative R");
  if (R == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(rois && J, NAWSOD_ERR_ARG, "roi_iou: null pointer");
  const size_t total = (size_t)R * R;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
  roi_iou_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(rois, R, J);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_cross_entropy_fwd(const float* X, const float* L, const float* Wt, int N, int C, int is_mean,
                                        float* Y, void* stream) {
  NAWSOD_REQUIRE(N > 0 && C > 0, NAWSOD_ERR_SHAPE, "cross_entropy: need N, C > 0");
  NAWSOD_REQUIRE((int64_t)N * C <= 4096, NAWSOD_ERR_UNSUPPORTED, "cross_entropy: N*C > 4096");
  NAWSOD_REQUIRE(X && L && Y, NAWSOD_ERR_ARG, "cross_entropy: null pointer");
  ce_fwd_kernel<<<1, 256, (size_t)N * C * sizeof(double), static_cast<cudaStream_t>(stream)>>>(X, L, Wt, N, C, is_mean, Y);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_cross_entropy_bwd(const float* X, const float* L, const float* Wt, const float* dY, int N, int C,
                                        int is_mean, float* dX, void* stream) {
  NAWSOD_REQUIRE(N > 0 && C > 0, NAWSOD_ERR_SHAPE, "cross_entropy_grad: need N, C > 0");
  NAWSOD_REQUIRE(X && L && dY && dX, NAWSOD_ERR_ARG, "cross_entropy_grad: null pointer");
  const int total = N * C;
  ce_bwd_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(X, L, Wt, dY, N, C, is_mean, dX);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}
