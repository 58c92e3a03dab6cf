/*
 * nawsod.h -- C ABI of libnawsod.so: the B200 (sm_100a) implementation of NA-fWebSOD's
 * per-proposal head.  This is the drop-in boundary: every entry
 * point replaces one Caffe2 operator (or one fused run of operators) of the reference and
 * keeps that operator's blob order and argument meaning.  Citations are relative to
 * /root/reference/detectron.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller owns all memory; the library never allocates device memory and keeps no
 *     hidden state between calls (the reference SGD op's iter_count_ is an explicit argument);
 *   - `stream` is a cudaStream_t passed as void*; calls are stream-ordered, re-entrant and
 *     never synchronise the device;
 *   - return value: NAWSOD_OK or an error code; nawsod_last_error() gives the message for
 *     the calling thread (the Python host raises RuntimeError, mirroring CAFFE_ENFORCE);
 *   - there is no CPU fallback: without a CUDA device every compute entry fails.
 */
#ifndef NAWSOD_H_
#define NAWSOD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  NAWSOD_OK = 0,
  NAWSOD_ERR_ARG = 1,         /* null pointer / bad enum / negative size      */
  NAWSOD_ERR_SHAPE = 2,       /* shape constraint violated                    */
  NAWSOD_ERR_ALIGN = 3,       /* pointer not 16-byte aligned where required   */
  NAWSOD_ERR_UNSUPPORTED = 4, /* combination not implemented                  */
  NAWSOD_ERR_CUDA = 5         /* CUDA runtime / launch failure                */
};

enum { NAWSOD_F32 = 0, NAWSOD_BF16 = 1 };

/* Feature-map / pooled-feature layouts.
 *   NCHW : the reference's layout.  X [N,C,H,W]; pooled Y [R,C,PH,PW] (FC K-index c*49+ph*7+pw).
 *   NHWC : channels-last.            X [N,H,W,C]; pooled Y [R,PH,PW,C] (FC K-index (ph*7+pw)*C+c). */
enum { NAWSOD_NCHW = 0, NAWSOD_NHWC = 1 };

const char* nawsod_last_error(void);
int nawsod_version(void);
/* Tuning knob for benchmarks ("pool_slab_bytes", "pool_chunks", ...); unknown keys fail. */
int nawsod_set_tuning(const char* key, int64_t value);

/* ---------------------------------------------------------------------------------------
 * Layout helpers (no reference counterpart: the reference is NCHW end to end).
 * in [B,rows,cols] -> out [B,cols,rows]; 4-byte elements (float / int32), optional
 * conversion of a float source to bf16 on the way out (out_dtype).
 * ------------------------------------------------------------------------------------- */
int nawsod_transpose_batched(const void* in, int64_t B, int64_t rows, int64_t cols,
                             void* out, int out_dtype, void* stream);

/* ---------------------------------------------------------------------------------------
 * a1 + a3: RoIPoolF([X, rois] -> [Y, argmax]; pooled_h, pooled_w, spatial_scale)
 *   replaces Caffe2 RoIPoolF as wired by modeling/detector.py:321-329 (arithmetic:
 *   ops/roi_loop_pool_op.cu:19-102 with RoIPoolF's deltas), optionally
 *   fused with RoIFeatureBoost([Y, S] -> Y) (ops/roi_feature_boost_op.cc:8-35,
 *   modeling/wsl_heads.py:668).
 *   X      : feature map, x_dtype / x_layout.
 *   rois   : [R,5] float (batch_idx, x1, y1, x2, y2) in network-input pixels.
 *   boost  : [R] float (obn_scores + 1) or NULL for plain RoIPoolF.
 *   Y      : pooled features, y_dtype / y_layout.  argmax: int32, same layout as Y, or NULL
 *            (is_test).  argmax is h*W+w inside the image plane, -1 for an empty bin.
 *   Bit-exact with the reference for x_dtype = y_dtype = F32 (any layout).
 * ------------------------------------------------------------------------------------- */
int nawsod_roi_pool_f_fwd(const void* X, int x_dtype, int x_layout, const float* rois,
                          const float* boost, int N, int C, int H, int W, int R,
                          float spatial_scale, int pooled_h, int pooled_w, void* Y,
                          int y_dtype, int y_layout, int32_t* argmax, void* stream);

/* a2 (+a3 gradient): RoIPoolFGradient([X, rois, argmax, dY] -> dX)
 *   (ops/roi_loop_pool_op.cu:105-140, zero fill :199-201; grad maker ops/roi_loop_pool_op.cc:85-96)
 *   optionally fused with RoIFeatureBoostGradient (ops/roi_feature_boost_op.cc:37-64): the
 *   incoming dY is multiplied by boost[r] first.  dX is float, zero-filled by the call. */
int nawsod_roi_pool_f_bwd(const void* dY, int dy_dtype, int y_layout, const int32_t* argmax,
                          const float* rois, const float* boost, int N, int C, int H, int W,
                          int R, int pooled_h, int pooled_w, float* dX, int dx_layout,
                          void* stream);

/* a3 stand-alone: RoIFeatureBoost([X, S] -> Y), in-place allowed (Y == X)
 *   (ops/roi_feature_boost_op.cc:8-35; its gradient is the same call on dY, :37-64). */
int nawsod_roi_feature_boost(const float* X, const float* S, int R, int64_t feature_size,
                             float* Y, void* stream);

/* ---------------------------------------------------------------------------------------
 * a4: FC stack GEMMs on the tcgen05 tensor cores (Caffe2 FC / Relu / Dropout and their
 *   gradient ops as wired by modeling/wsl_heads.py:674-679 and modeling/webly_heads.py:490-498).
 *   Every matrix is row-major with an explicit leading dimension (elements between rows), so
 *   column slices of a wider buffer (e.g. one stack's half of a fused [R, 8192] activation)
 *   are addressed without copies.  ab_dtype: NAWSOD_BF16 (bf16 operands, fp32 accumulate) or
 *   NAWSOD_F32 (operands read as TF32, fp32 accumulate).  Operand base pointers must be
 *   16-byte aligned and leading dimensions multiples of 16 bytes (TMA).
 *
 * nawsod_fc_fwd   FC([A, W, b] -> Y) [+ Relu] [+ Dropout]:
 *     Y[M,N] = epilogue( A[M,K] . W[N,K]^T + bias[N] );  bias may be NULL.
 *     NAWSOD_FC_RELU: max(.,0).  NAWSOD_FC_DROPOUT: Y *= 2 * keep[M,N] (Caffe2 Dropout ratio 0.5,
 *     scale 1/(1-ratio)); keep = mask (uint8 0/1) when mask != NULL (injected, for parity runs),
 *     else a counter-based hash of (dropout_seed, m, n) when dropout_seed != 0, else 1.
 * nawsod_fc_bwd_x FCGradient's dX (+ the ReluGradient / DropoutGradient of the layer below):
 *     dA[M,K] = dY[M,N] . W[N,K];  NAWSOD_FC_RELU: dA *= (act_below[M,K] > 0);
 *     NAWSOD_FC_DROPOUT: dA *= 2 (and *= mask_below[M,K] when mask_below != NULL; with
 *     act_below = the post-dropout activation the mask is implied by act_below > 0).
 * nawsod_fc_bwd_w FCGradient's dW, db:
 *     dW[N,K] (float) = dY[M,N]^T . A[M,K];  db[N] (float) = sum_m dY[m,:] (db may be NULL).
 *     NAWSOD_FC_ACCUMULATE: add into dW / db instead of overwriting.
 * NAWSOD_FC_ACCUMULATE on fwd / bwd_x (float outputs only): the prior contents of Y / dA are added to the
 *     product BEFORE bias, activation, dropout and their gradients -- how the split-operand fp32 path
 *     (nawsod_split_tf32 below) sums its three tensor-core passes x.y = lo.hi' + hi.lo' + hi.hi'.
 * NAWSOD_FC_ROUND_TF32 (fwd / bwd_x, float outputs): round the stored values to the nearest
 *     TF32 so that the next GEMM's tensor-core truncation is exact (keeps the fp32 path unbiased).
 * ------------------------------------------------------------------------------------- */
enum { NAWSOD_FC_RELU = 1, NAWSOD_FC_DROPOUT = 2, NAWSOD_FC_ACCUMULATE = 4, NAWSOD_FC_ROUND_TF32 = 8 };

int nawsod_fc_fwd(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                  const uint8_t* mask, int64_t ldmask, uint64_t dropout_seed, int M, int N, int K,
                  int ab_dtype, void* Y, int64_t ldy, int y_dtype, int flags, void* stream);

/* nawsod_fc_fwd whose weight reads are GATED on the data-parallel peer exchange (a11): rows
 *   [g * gate_rows, (g + 1) * gate_rows) of W are read only once the gate_nflags device words
 *   gate_flags[g * gate_nflags ...] have reached gate_seq (the "operands of exchange bucket g have
 *   landed" flags of nawsod_p2p_scatter / nawsod_p2p_signal), g < gate_groups; gate_rows a multiple of
 *   256.  The kernel starts on the row groups that are there and meets the others as they arrive, so
 *   the tail of one step's exchange hides behind the next step's fc6.  A flag that does not arrive
 *   within gate_timeout_ms sets *gate_status (device uint32, may be NULL) and the kernel proceeds. */
int nawsod_fc_fwd_gated(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                        const uint8_t* mask, int64_t ldmask, uint64_t dropout_seed, int M, int N,
                        int K, int ab_dtype, void* Y, int64_t ldy, int y_dtype, int flags,
                        const void* gate_flags, int gate_groups, int gate_nflags, int gate_rows,
                        uint32_t gate_seq, int64_t gate_timeout_ms, void* gate_status, void* stream);

int nawsod_fc_bwd_x(const void* dY, int64_t lddy, const void* W, int64_t ldw,
                    const void* act_below, int64_t ldact, int act_dtype,
                    const uint8_t* mask_below, int64_t ldmask, int M, int N, int K,
                    int ab_dtype, void* dA, int64_t ldda, int da_dtype, int flags,
                    void* stream);

int nawsod_fc_bwd_w(const void* dY, int64_t lddy, const void* A, int64_t lda, int M, int N,
                    int K, int ab_dtype, float* dW, int64_t lddw, float* db, int flags,
                    void* stream);

/* The same three GEMMs over S independent problems of one shape in ONE launch (the head's clean and
 * noisy stacks, webly_heads.py:490-498: twice the tiles per launch, half the launches).  Stack s uses
 * operand + s * stride (strides in elements of the operand's type; bias / db in floats; masks in
 * bytes) and draws seeded dropout bits from dropout_seed + s.  S = 1 ignores the strides. */
int nawsod_fc_fwd_stacks(const void* A, int64_t lda, int64_t sA, const void* W, int64_t ldw,
                         int64_t sW, const float* bias, int64_t sbias, const uint8_t* mask,
                         int64_t ldmask, int64_t smask, uint64_t dropout_seed, int S, int M, int N,
                         int K, int ab_dtype, void* Y, int64_t ldy, int64_t sY, int y_dtype,
                         int flags, void* stream);
int nawsod_fc_bwd_x_stacks(const void* dY, int64_t lddy, int64_t sdY, const void* W, int64_t ldw,
                           int64_t sW, const void* act_below, int64_t ldact, int64_t sact,
                           int act_dtype, const uint8_t* mask_below, int64_t ldmask, int64_t smask,
                           int S, int M, int N, int K, int ab_dtype, void* dA, int64_t ldda,
                           int64_t sdA, int da_dtype, int flags, void* stream);
int nawsod_fc_bwd_w_stacks(const void* dY, int64_t lddy, int64_t sdY, const void* A, int64_t lda,
                           int64_t sA, int S, int M, int N, int K, int ab_dtype, float* dW,
                           int64_t lddw, int64_t sdW, float* db, int64_t sdb, int flags,
                           void* stream);

/* FCGradient's db alone: db[s][N] (float) = column sums of dY[s][M,N] -- what nawsod_fc_bwd_w[_stacks] computes when db is
 * given, as a call of its own (FCGradient's db output without its dW). */
int nawsod_fc_bias_grad(const void* dY, int64_t lddy, int64_t sdY, int S, int M, int N, int ab_dtype,
                        float* db, int64_t sdb, int flags, void* stream);

/* Operand staging for the GEMMs above: float -> bf16, and float -> nearest-TF32 (kept in a
 * float container; dst may alias src) of a [rows, cols] matrix. */
int nawsod_convert_f32_to_bf16(const float* src, int64_t ld_src, int64_t rows, int64_t cols,
                               void* dst, int64_t ld_dst, void* stream);
int nawsod_round_to_tf32(const float* src, int64_t ld_src, int64_t rows, int64_t cols,
                         float* dst, int64_t ld_dst, void* stream);
/* fp32 path at the reference's precision (Caffe2 FC = cuBLAS sgemm, modeling/wsl_heads.py:674-679) on the TF32
 * tensor cores: src = hi + lo with hi = nearest TF32 of src (hi may alias src) and lo = nearest TF32 of the exact
 * remainder (lo must not alias).  Three passes hi.hi' + lo.hi' + hi.lo' (NAWSOD_FC_ACCUMULATE) reproduce the fp32
 * product to ~2^-21 relative; one pass (the TF32 path) stops at ~2^-11 per operand. */
int nawsod_split_tf32(const float* src, int64_t ld_src, int64_t rows, int64_t cols, float* hi,
                      int64_t ld_hi, float* lo, int64_t ld_lo, void* stream);

/* ---------------------------------------------------------------------------------------
 * a5..a9: the two-stream MIL head, noise-aware class weights, weighted multi-label CE and
 * the whole backward to the fc8 logits, as ONE fused kernel
 *   (modeling/wsl_heads.py:49-55,213-227; modeling/webly_heads.py:32-74,123-197,265-391;
 *    ops/roi_iou_op.cu:28-62; ops/cross_entropy_wsl_op.cc:88-180).
 *   fc8c, fc8d          : [R,C] float logits of the clean stack, row pitch ld_logits elements
 *                         (fc8c | fc8d may be the two halves of one fused [R,2C] FC output).
 *   nfc8c, nfc8d        : [R,C] float logits of the noisy stack (same pitch), or both NULL (plain WSDDN).
 *   rois                : [R,5]; rois of image b are rows roi_offsets_host[b]..[b+1]
 *                         (contiguous per image, the order contract of ops/roi_score_reshape_op.cc:30-44).
 *   roi_offsets         : [B+1] int32 DEVICE array.
 *   labels_oh           : [B,C] float (may be soft: mixup).
 *   flags               : NAWSOD_MIL_ENTROPY (WEBLY.ENTROPY), NAWSOD_MIL_MEAN (WSL.MEAN_LOSS),
 *                         NAWSOD_MIL_BACKWARD (also produce d_*).
 * outputs (any may be NULL except loss when BACKWARD is off):
 *   rois_pred[R,C], cls_prob[B,C], rois_pred_noise[R,C], cls_prob_noise[B,C],
 *   class_weight[B,C], class_weight_noise[B,C], loss[B,2] (loss_cls, loss_cls_noise per image),
 *   d_fc8c, d_fc8d, d_nfc8c, d_nfc8d [R,C], row pitch ld_grads (loss-gradient seed 1.0 each,
 *   utils/blob.py:167-173).
 *   workspace: device scratch >= nawsod_mil_workspace_bytes(R, C, B).
 * ------------------------------------------------------------------------------------- */
enum { NAWSOD_MIL_ENTROPY = 1, NAWSOD_MIL_MEAN = 2, NAWSOD_MIL_BACKWARD = 4 };

int64_t nawsod_mil_workspace_bytes(int R, int C, int B);

int nawsod_mil_head_fwd_bwd(const float* fc8c, const float* fc8d, const float* nfc8c,
                            const float* nfc8d, int64_t ld_logits, const float* rois,
                            const int32_t* roi_offsets, const float* labels_oh, int R, int C,
                            int B, int flags,
                            float* rois_pred, float* cls_prob, float* rois_pred_noise,
                            float* cls_prob_noise, float* class_weight,
                            float* class_weight_noise, float* loss, float* d_fc8c,
                            float* d_fc8d, float* d_nfc8c, float* d_nfc8d, int64_t ld_grads,
                            void* workspace, void* stream);

/* a7 stand-alone: RoIIoU([rois] -> [J]) (ops/roi_iou_op.cc:11-18, ops/roi_iou_op.cu:28-84). J [R,R]. */
int nawsod_roi_iou(const float* rois, int R, float* J, void* stream);

/* a8 stand-alone: [Weighted]CrossEntropyWithLogits([X, L(, W)] -> [Y]; is_mean) and gradient
 *   ([X, L(, W), dY] -> [dX]) (ops/cross_entropy_wsl_op.cc:8-180).  X, L, W: [N,C]; W may be NULL. */
int nawsod_cross_entropy_fwd(const float* X, const float* L, const float* Wt, int N, int C,
                             int is_mean, float* Y, void* stream);
int nawsod_cross_entropy_bwd(const float* X, const float* L, const float* Wt, const float* dY,
                             int N, int C, int is_mean, float* dX, void* stream);

/* ---------------------------------------------------------------------------------------
 * a10: ACMWeightDecayMomentumSGDUpdate([g, m, lr, p, acc] -> [g, m, p, acc])
 *   (ops/acm_weightdecay_momentum_sgd_op.h:48-112, wired by modeling/optimizer_wsl.py:96-137),
 *   fused into one pass.  In place on m, p, acc.  `lr` is a 1-element DEVICE float
 *   (the reference's `lr` blob).  iter_count = number of calls already made on this
 *   parameter (call 0 zero-initialises m and acc, .h:62-69).  p_shadow (optional) receives
 *   the updated parameter as the tensor-core GEMM operand: bf16 (shadow_dtype NAWSOD_BF16) or
 *   float rounded to the nearest TF32 (shadow_dtype NAWSOD_F32).
 * ------------------------------------------------------------------------------------- */
int nawsod_sgd_update(const float* g, float* m, const float* lr, float* p, float* acc,
                      int64_t n, float momentum, float weight_decay, float lr_mult,
                      int iter_size, int gpu_num, int64_t iter_count, void* p_shadow,
                      int shadow_dtype, void* stream);

/* a10 + a11 on the owner rank of a parameter slice: the gradient is the sum, in the order given, of
 *   n_grads contributions (the rank's own slice and the copies its peers deposited), followed by the
 *   same update as nawsod_sgd_update with iter_size 1 (the reference all-reduces, then updates:
 *   modeling/optimizer_wsl.py:52-72, 96-137).  abort_flag (optional, device uint32): when the word is
 *   non-zero at launch time -- the status word of a nawsod_p2p_wait whose watchdog fired -- the call
 *   leaves m, p and the shadow untouched instead of updating from an incomplete sum. */
int nawsod_sgd_update_reduce(const float* const* grads, int n_grads, float* m, const float* lr,
                             float* p, int64_t n, float momentum, float weight_decay,
                             float lr_mult, int gpu_num, int64_t iter_count, void* p_shadow,
                             int shadow_dtype, const void* abort_flag, void* stream);

/* ---------------------------------------------------------------------------------------
 * a11: peer-to-peer plumbing of the gradient exchange (replaces the NCCLAllreduce ops of
 *   modeling/optimizer_wsl.py:52-72 when the ranks of one NVSwitch box map each other's buffers).
 *   nawsod_p2p_copy: cudaMemcpyAsync (copy engine) between local / peer-mapped device buffers.
 *   nawsod_p2p_signal: store `value` (system-scope release) into each of n flag words, in stream
 *   order after the copies that precede it.  nawsod_p2p_wait: block the stream until all n flag
 *   words (device uint32, contiguous) have reached `value` (wrap-safe); after timeout_ms the
 *   wait gives up and sets *status (device uint32, may be NULL) to 1 instead of hanging.
 *   nawsod_p2p_enable_peer_access: let kernels and copies of the current device address memory of
 *   `peer_device` (cudaDeviceEnablePeerAccess; idempotent).
 * ------------------------------------------------------------------------------------- */
int nawsod_p2p_enable_peer_access(int peer_device);
/* export: the cudaIpcMemHandle_t (64 bytes) of the allocation holding `ptr` and ptr's offset inside it */
int nawsod_p2p_get_mem_handle(const void* ptr, void* handle_out, int64_t handle_bytes,
                              int64_t* offset_out);
/* map a peer process's allocation (its cudaIpcMemHandle_t, 64 bytes) into the current device's context */
int nawsod_p2p_open_mem_handle(const void* handle, int64_t handle_bytes, void** base);
int nawsod_p2p_close_mem_handle(void* base);
int nawsod_p2p_copy(void* dst, const void* src, int64_t bytes, void* stream);
int nawsod_p2p_signal(void* const* flag_ptrs, int n, uint32_t value, void* stream);
/* SM-driven scatter (srcs differ) / broadcast (srcs equal): one launch copies `bytes` from srcs[i] to
 * dsts[i] (local or peer-mapped) for each of npeers destinations with 16-byte posted stores, then
 * stores `value` (system-scope release) into the nflags flag words.  No shared memory: its CTAs
 * co-reside with the persistent GEMMs.  `slot` in [0,128) names the completion counter (distinct
 * for launches that may overlap in time). */
int nawsod_p2p_scatter(const void* const* srcs, void* const* dsts, int npeers, int64_t bytes,
                       void* const* flag_ptrs, int nflags, uint32_t value, int slot, void* stream);
int nawsod_p2p_wait(const void* flags, int n, uint32_t value, int64_t timeout_ms, void* status,
                    void* stream);

/* ---------------------------------------------------------------------------------------
 * SURVEY.md 8f "next" rows: the test-time wrapper around the head.
 *
 * N1  core/test_wsl.py:100-178 (im_detect_bbox), :181-281 (im_detect_bbox_aug), :998-1059.
 *   nawsod_project_rois   rois[R,5] = [batch_idx, (float)((double)box * im_scale)]; flip_width >= 0
 *                         first flips the boxes horizontally (utils/boxes.py:246-251:
 *                         x1' = W - x2 - 1, x2' = W - x1 - 1); flip_width < 0: no flip.  obn_out = obn_scores + 1
 *                         (core/test_wsl.py:1058) when both are given.
 *   nawsod_dedup_rois     hashes = round(rois * DEDUP_BOXES) . [1,1e3,1e6,1e9,1e12];
 *                         np.unique(hashes, return_index, return_inverse) (core/test_wsl.py:125-133):
 *                         index[u] = first row with the u-th smallest hash (entries u >= num_unique
 *                         repeat index[0]), inv_index[r] = rank of row r's hash, num_unique[0].
 *                         R <= 16384 (the shipped TEST.PROPOSAL_LIMIT is 9999).  All outputs are device arrays; roi_offsets (or NULL) receives
 *                         {0, num_unique}, the row range nawsod_mil_head_fwd_bwd takes for one image, so the
 *                         head can run on the unique set without a host round trip.
 *   nawsod_gather_rows    dst[i,:] = src[index[i],:]   (rois / obn_scores / boxes of the unique set)
 *   nawsod_scatter_scores cls_prob[r,:] (=|+=) concat(rois_pred[u,:1], rois_pred[u,:]), u = inv_index[r]
 *                         (modeling/wsl_heads.py:57-67 + core/test_wsl.py:173-176); inv_index NULL = identity;
 *                         accumulate != 0 adds in float32 (np.mean(scores_ts, axis=0) sums the passes
 *                         in order, core/test_wsl.py:262-263) and nawsod_scores_finalize divides by T.
 * ------------------------------------------------------------------------------------- */
int nawsod_project_rois(const float* boxes, int R, double im_scale, double flip_width, int batch_idx,
                        float* rois, const float* obn_scores, float* obn_out, void* stream);
int nawsod_dedup_rois(const float* rois, int R, float dedup_scale, int32_t* index, int32_t* inv_index,
                      int32_t* num_unique, int32_t* roi_offsets, void* stream);
int nawsod_gather_rows(const float* src, const int32_t* index, int n, int cols, float* dst, void* stream);
int nawsod_scatter_scores(const float* rois_pred, int64_t ld, const int32_t* inv_index, int R, int C,
                          int accumulate, float* cls_prob, void* stream);
int nawsod_scores_finalize(float* acc, int64_t n, int count, void* stream);

/* N2  box_results_with_nms_and_limit (core/test_wsl.py:803-863) with the greedy NMS of
 *   utils/cython_nms.pyx:38-93 (float32 arithmetic, suppress when ovr >= nms_thresh), classes
 *   1..num_classes-1; scores [R,num_classes], boxes [R,4] (COORD_HEUR 'ID': one box per proposal).
 *   keep [num_classes,R] uint8, num_keep [num_classes] int32, image_thresh [1] float or NULL
 *   (-FLT_MAX when the detections_per_im limit did not bite).  Equal scores are visited higher row
 *   first.  R <= 16384. */
int nawsod_nms_and_limit(const float* scores, const float* boxes, int R, int num_classes,
                         float score_thresh, float nms_thresh, int detections_per_im, uint8_t* keep,
                         int32_t* num_keep, float* image_thresh, void* stream);

/* N4  MinEntropyLoss([X, L] -> Y) / MinEntropyLossGradient([X, L, dY] -> dX)
 *   (ops/min_entropy_loss_op.cu:34-66,70-152): X [N,C] probabilities, L [1,C] labels (B must be 1),
 *   Y scalar = sum_{L[c]>=0.5} -p log p / (1 + count); norm / norm_ws: one float of scratch. */
int nawsod_min_entropy_loss_fwd(const float* X, const float* L, int N, int C, int B, float* Y,
                                float* norm, void* stream);
int nawsod_min_entropy_loss_bwd(const float* X, const float* L, const float* dY, int N, int C, int B,
                                float* dX, float* norm_ws, void* stream);

/* N3  the training-input contract: proposals -> the `rois` / `obn_scores` / `labels_oh` blobs, crop offsets,
 *   bagging-mixup (roi_data/wsl.py:87-225, roi_data/loader_wsl.py:149-168, tools/convert_mcg.py:37-49).
 *   nawsod_sample_rois   _sample_rois + _project_im_rois (roi_data/wsl.py:101-111,212-225) for ONE image: boxes
 *                        [R,4] float (the first min(BATCH_SIZE_PER_IM, n) rows of roidb['boxes'], original image
 *                        pixels) are clipped to the crop window (x1,y1,x2,y2; minibatch_wsl.py:63-64), shifted by
 *                        its origin and scaled in double; rois[R,5] = (batch_idx, x1, y1, x2, y2) float;
 *                        obn_out = obn_scores + 1 (both or neither).  rois may point into a larger [sum R,5] blob
 *                        (add_wsl_blobs concatenates per-image blobs, roi_data/wsl.py:59-85).
 *   nawsod_image_labels  labels_oh[num_classes-1] one-hot union of the rows with gt_classes > 0; labels_int32[1] =
 *                        class - 1 of the LAST such row (roi_data/wsl.py:139-155), -1 if there is none
 *                        (the reference asserts); labels_int32 may be NULL.
 *   nawsod_bagging_mixup out = lam0 * x0 + lam1 * x1 in float32, each product and sum rounded separately
 *                        (loader_wsl.py:158-164, `data` and `labels_oh` blobs); out may alias x0 or x1.
 *   nawsod_set_column    a[:, col] = value for a [rows, ld] float matrix (`blobs['rois'][:, 0] = 0`, :165).
 *   nawsod_convert_mcg_boxes  1-indexed (y1,x1,y2,x2) double -> 0-indexed (x1,y1,x2,y2) uint16 with the script's
 *                        uint16 arithmetic (tools/convert_mcg.py:45-49). */
int nawsod_sample_rois(const float* boxes, int R, double im_scale, int crop_x1, int crop_y1, int crop_x2,
                       int crop_y2, int batch_idx, float* rois, const float* obn_scores, float* obn_out,
                       void* stream);
int nawsod_image_labels(const int32_t* gt_classes, int n, int num_classes, float* labels_oh,
                        int32_t* labels_int32, void* stream);
int nawsod_bagging_mixup(const float* x0, const float* x1, int64_t n, float lam0, float lam1, float* out,
                         void* stream);
int nawsod_set_column(float* a, int rows, int64_t ld, int col, float value, void* stream);
int nawsod_convert_mcg_boxes(const double* bboxes, int R, uint16_t* boxes_out, void* stream);

/* N4 (second half)  the frozen VGG16 conv body in channels-last bf16 (detectron/modeling/VGG16.py:9-58; Caffe2 Conv / Relu /
 *   MaxPool, forward only: TRAIN.FREEZE_CONV_BODY).  EXPERIMENTAL: compiled, not yet run on hardware.  A 3x3 convolution
 *   is nawsod_im2col3x3 followed by nawsod_fc_fwd with NAWSOD_FC_RELU on the patch matrix (W permuted to [Cout, (kh,kw,c)]).
 *   nawsod_im2col3x3   X [N,H,W,C] bf16 -> cols [N*H*W, 9*C] bf16, K-order (kh, kw, c); stride 1, pad = dilation
 *                      (1 or 2, the combinations VGG16.py:10-56 uses), zeros outside the image; C % 8 == 0.
 *   nawsod_maxpool2x2  MaxPool(kernel=2, pad=0, stride) on [N,H,W,C] bf16 -> [N,(H-2)/stride+1,(W-2)/stride+1,C];
 *                      stride 2 (pool1-3, pool4 of the 1/16 body) or 1 (pool4 with WSL.DILATION 2, VGG16.py:40-41). */
/* Conv(3x3, stride 1, pad = dilation) + bias [+ Relu] as an implicit GEMM on the tcgen05 tensor cores: X [N,H,W,Cin] bf16
 *   channels-last, Wmat [Cout, 9*Cin] bf16 with K-order (kh, kw, c), bias [Cout] float, Y [N,H,W,Cout] bf16.  The A operand
 *   of a k-block is a shifted [8, 16, 64-channel] box of X loaded by ONE 4-D tiled TMA copy (out-of-bounds = the zero
 *   padding): no patch matrix.  Cin % 64 == 0, Cout % 32 == 0, dilation 1 or 2.  Bit-identical to
 *   nawsod_im2col3x3 + nawsod_fc_fwd. */
int nawsod_conv3x3_relu(const void* X, int N, int H, int W, int Cin, const void* Wmat, const float* bias,
                        int Cout, int dilation, int relu, void* Y, void* stream);
int nawsod_im2col3x3(const void* X, int N, int H, int W, int C, int dilation, void* cols, void* stream);
int nawsod_maxpool2x2(const void* X, int N, int H, int W, int C, int stride, void* Y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NAWSOD_H_ */
