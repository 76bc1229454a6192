/*
 * soswsod_b200.h -- C ABI of libsoswsod_b200.so: the B200 (sm_100a) implementation of the SoS-WSOD
 * Stage-1 OICR+ ROI-head hot path.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer unless it says "host"; the caller (PyTorch shim) owns and
 *     allocates every buffer including workspaces -- the library never allocates device memory,
 *     never synchronises and never changes the current device;
 *   - every entry point is stream-ordered on `stream` (a cudaStream_t passed as void*) and re-entrant;
 *   - return value: 0 on success, negative SOSWSOD_ERR_* otherwise; soswsod_last_error() returns a
 *     thread-local human-readable message.  Nothing ever calls exit().
 *   - matrices are row-major; `ld*` are leading dimensions in ELEMENTS.
 *
 * Reference interfaces these entry points replace (paths under /root/reference/uwsod/):
 *   W/ = projects/WSL/wsl/   D/ = detectron2/
 */
#ifndef SOSWSOD_B200_H_
#define SOSWSOD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOSWSOD_ABI_VERSION 10

#define SOSWSOD_OK 0
#define SOSWSOD_ERR_INVALID (-1)     /* bad shape / null pointer / misalignment */
#define SOSWSOD_ERR_CUDA (-2)        /* a CUDA runtime / driver call failed */
#define SOSWSOD_ERR_UNSUPPORTED (-3) /* size outside what the kernels handle */
#define SOSWSOD_ERR_WORKSPACE (-4)   /* workspace too small */

#define SOSWSOD_DTYPE_F32 0
#define SOSWSOD_DTYPE_BF16 1
#define SOSWSOD_ARGMAX_I32 0
#define SOSWSOD_ARGMAX_U16 1 /* 0xFFFF encodes -1; requires h*w < 65535 */

typedef void* soswsod_stream_t;

int soswsod_abi_version(void);
const char* soswsod_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * (1) ROI max-pool.  Replaces torchvision `roi_pool` as called by ROIPooler
 *     (W/modeling/poolers.py:183-186, 263-270) fused with the objectness scaling of
 *     W/modeling/roi_heads/roi_heads_oicrplus.py:200-221; operator pattern of
 *     W/layers/roi_loop_pool.py:9-35 (`_C.roi_loop_pool_forward/_backward`).
 *
 *   feat       fp32 [n, c, h, w] (NCHW)
 *   rois       fp32 [num_rois, 5] = (batch index, x1, y1, x2, y2)
 *   out_f32    fp32 [num_rois, c, ph, pw]  raw pooled maxima, bit-equal to torchvision (may be NULL)
 *   argmax     [num_rois, c, ph, pw] index h*W+w inside the (n,c) plane, -1 for an empty bin;
 *              int32 or uint16 per `argmax_dtype` (may be NULL)
 *   out_bf16   bf16 [num_rois, ld_bf16] = pooled * (row_scale[r] + row_scale_bias), the fc6 GEMM
 *              operand (may be NULL).  row_scale NULL => factor 1.
 *   plan       optional (NULL allowed): the per-call plan written by soswsod_roi_pool_plan for the SAME
 *              rois / n / h / w / spatial_scale / row_scale (7x7 bins only).  With a plan, uint16
 *              arg-max and only out_bf16 requested (the head engine's operand mode) the forward runs
 *              the channel-interleaved window-table kernel; the backward (bf16 or fp32 grad, uint16
 *              arg-max) runs the two-planes-per-warp kernel.  Results are identical with and without.
 * ------------------------------------------------------------------------------------------- */
/* Bytes of a plan for num_rois rois (0 when the pooled size has no plan). */
size_t soswsod_roi_pool_plan_bytes(int num_rois, int pooled_h, int pooled_w);
/* Groups the rois by image and stores every roi's bin bounds, scale factor and backward colouring
 * (bin arithmetic of torchvision roi_pool, done once per roi).  plan: 128-byte aligned, n <= 64. */
int soswsod_roi_pool_plan(const float* rois, int num_rois, int n, int h, int w, int pooled_h, int pooled_w,
                          float spatial_scale, const float* row_scale, float row_scale_bias, void* plan,
                          size_t plan_bytes, soswsod_stream_t stream);
int soswsod_roi_pool_forward(const float* feat, int n, int c, int h, int w, const float* rois, int num_rois,
                             int pooled_h, int pooled_w, float spatial_scale, const float* row_scale,
                             float row_scale_bias, float* out_f32, void* argmax, int argmax_dtype,
                             void* out_bf16, long long ld_bf16, const void* plan, size_t plan_bytes,
                             soswsod_stream_t stream);

/* Atomic-free backward: grad_feat[n,c,h,w] (fully overwritten) = sum over (roi,bin) with
 * argmax == (h,w) of grad_out[roi, c*ph*pw + bin] * (row_scale[roi] + row_scale_bias).
 * grad_out is [num_rois, ld_grad] fp32 or bf16.  `argmax` must be the tensor soswsod_roi_pool_forward
 * wrote for the same rois / spatial_scale (the kernel relies on arg-max cells lying inside their
 * bins to schedule conflict-free updates).  Replaces torchvision's atomicAdd backward (same call
 * sites as above). */
int soswsod_roi_pool_backward(const void* grad_out, int grad_dtype, long long ld_grad, const void* argmax,
                              int argmax_dtype, const float* rois, int num_rois, const float* row_scale,
                              float row_scale_bias, int n, int c, int h, int w, int pooled_h, int pooled_w,
                              float spatial_scale, float* grad_feat, const void* plan, size_t plan_bytes,
                              soswsod_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (2) fc6/fc7/head GEMMs on tcgen05 tensor cores, fp32 accumulation in TMEM, TMA-fed.
 *     Replaces nn.Linear + F.relu_ + F.dropout in DiscriminativeAdaptionNeck.forward
 *     (W/modeling/roi_heads/box_head.py:82-91), the cls/det Linear of WSDDNOutputLayers
 *     (W/modeling/roi_heads/fast_rcnn_wsddn.py:558-559), cls_score/bbox_pred of OICROutputLayers
 *     (W/modeling/roi_heads/fast_rcnn_oicr.py:517-519) and their autograd backward.
 *
 *   D[M,N] = epilogue( sum_k A[m,k] * B[n,k] )
 *   A: bf16, K-major  [M, lda]  (a_mn_major = 0)  or MN-major [K, lda] (a_mn_major = 1)
 *   B: bf16, K-major  [N, ldb]  (b_mn_major = 0)  or MN-major [K, ldb] (b_mn_major = 1)
 *   epilogue, in order: + bias[n] (if bias) ; relu (if relu) ;
 *                       * (mask_src[m,n] > 0 ? mask_scale : 0) (if mask_src, bf16 [M, ld_mask]) ;
 *                       dropout keep/scale with a counter-based hash of (seed, m, n) (if dropout_p > 0)
 *   D: fp32 or bf16 per d_dtype.  lda/ldb multiples of 8 elements, bases 16-byte aligned.
 *   sched_workspace: NULL (tiles statically interleaved over the persistent CTAs) or 8 bytes of device
 *     memory, zeroed ONCE by the caller and reused by every launch on the same stream: the CTAs then take
 *     tiles from a global counter (a CTA slowed by a co-resident kernel takes fewer), and the last CTA
 *     re-arms the counter.  Launches that may run concurrently need separate workspaces.
 * ------------------------------------------------------------------------------------------- */
int soswsod_gemm_bf16(const void* a, long long lda, int a_mn_major, const void* b, long long ldb,
                      int b_mn_major, void* d, long long ldd, int d_dtype, int m, int n, int k,
                      const float* bias, int relu, const void* mask_src, long long ld_mask, float mask_scale,
                      float dropout_p, unsigned long long dropout_seed, void* sched_workspace,
                      soswsod_stream_t stream);

/* The dropout keep-mask the GEMM epilogue applies, materialised as uint8 [m, n] (1 = keep): lets a
 * test feed the identical mask to the oracle. */
int soswsod_dropout_mask(unsigned char* mask, int m, int n, float dropout_p, unsigned long long seed,
                         soswsod_stream_t stream);

/* out[r, c] = bf16(in[r, c] * col_scale[c] * m[r, c]) and (optionally) out_t[c, r] = same, in fp32 [rows, ld_in].
 * m = 1 without mask_src; with it (bf16 [rows, ld_mask], the saved post-ReLU(+dropout) activation)
 * m = mask_scale where mask_src > 0 and 0 elsewhere -- the backward of F.relu_ + F.dropout of
 * W/modeling/roi_heads/box_head.py:88-90 applied to an incoming gradient.  col_scale / mask_src may be NULL;
 * out or out_t may be NULL. */
int soswsod_cast_f32_bf16(const float* in, long long ld_in, int rows, int cols, const float* col_scale,
                          void* out, long long ld_out, void* out_t, long long ld_out_t,
                          const void* mask_src, long long ld_mask, float mask_scale, soswsod_stream_t stream);
/* out_t[c, r] = in[r, c] for bf16 [rows, ld_in]. */
int soswsod_transpose_bf16(const void* in, long long ld_in, int rows, int cols, void* out_t,
                           long long ld_out_t, soswsod_stream_t stream);
/* out[c] = sum_r in[r, c] (bias gradients), fixed summation order; in bf16 or fp32 per in_dtype.
 * With a workspace of soswsod_colsum_workspace_bytes() (16-byte aligned; NULL allowed) the matrix is
 * streamed once with 16-byte loads into per-chunk partial sums that a second launch adds in chunk order. */
size_t soswsod_colsum_workspace_bytes(int rows, int cols, int in_dtype);
int soswsod_colsum(const void* in, int in_dtype, long long ld_in, int rows, int cols, float* out,
                   void* workspace, size_t workspace_bytes, soswsod_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (3) Fused WSDDN two-stream head: scores, image scores, BCE and the gradient w.r.t. both logit
 *     blocks in one launch (one 8-CTA cluster per view, reductions through distributed shared memory).  Replaces WSDDNOutputLayers.forward softmax product
 *     (W/modeling/roi_heads/fast_rcnn_wsddn.py:566-567), WSDDNOutputs.predict_probs_img (:360-375) and
 *     binary_cross_entropy_loss (:340-358) with MEAN_LOSS.
 *
 *   logits   fp32 [num_views*R, ld]; view v = rows [v*R, (v+1)*R); cls logits in columns
 *            [col_cls, col_cls+C), det logits in [col_det, col_det+C)
 *   gt_onehot fp32 [C]
 *   scores   fp32 [num_views, R, C]            (softmax_c(cls) * softmax_r(det))
 *   img_scores fp32 [num_views, C]             (clamped to [1e-6, 1-1e-6])
 *   loss     fp32 [num_views]                  (BCE mean over C)
 *   dlogits  fp32 [num_views*R, ld_d] or NULL: d loss_v / d logits written into the same two column
 *            blocks (unit upstream gradient).
 * ------------------------------------------------------------------------------------------- */
int soswsod_wsddn_forward(const float* logits, long long ld, int col_cls, int col_det, int num_views, int R,
                          int C, const float* gt_onehot, float* scores, float* img_scores, float* loss,
                          float* dlogits, long long ld_d, soswsod_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (4) OICR refinement branches.
 *
 * soswsod_oicr_avg_scores: the view-averaged, detached scores each branch mines its seeds from
 *   (W/modeling/roi_heads/roi_heads_oicrplus.py:290-294 and :390-395):
 *   prev[0] = ((s_0 + s_1) + s_2 + s_3)/4 of the WSDDN scores [V,R,C] (column C zero-filled),
 *   prev[k] = mean_v softmax(logits_v[:, col_ref0 + (k-1)*ref_stride : +C+1]) for k >= 1.
 *   prev is fp32 [K, R, C+1].
 * ------------------------------------------------------------------------------------------- */
int soswsod_oicr_avg_scores(const float* wsddn_scores, const float* logits, long long ld, int col_ref0,
                            int ref_stride, int num_views, int R, int C, int K, float* prev,
                            soswsod_stream_t stream);

/* soswsod_image_level_gt: get_image_level_gt (W/modeling/roi_heads/roi_heads.py:144-164) on the device:
 *   gt_classes int64 / int32 [n] (any order, duplicates allowed) -> gt_list int32 [C] = the distinct classes
 *   ascending, padded with -1; gt_count int32 [1]; gt_onehot fp32 [C].  C <= 128.  No host synchronisation
 *   (torch.unique, which the reference calls, reads its output size back). */
int soswsod_image_level_gt(const void* gt_classes, int is_int64, int n, int C, int32_t* gt_list, int32_t* gt_count,
                           float* gt_onehot, soswsod_stream_t stream);

/* soswsod_oicr_mine_label: pseudo-GT mining + proposal labelling for all K branches.
 *   Replaces get_pgt_top_k / get_pgt_mist (roi_heads_oicrplus.py:559-757: per-GT-class top-k,
 *   threshold with rank 0 forced, class-agnostic NMS at `nms_thr`), pairwise_iou
 *   (D/structures/boxes.py:329-361), Matcher (D/modeling/matcher.py:63-111) and
 *   label_and_sample_proposals (W/modeling/roi_heads/roi_heads.py:266-375).
 *   Ordering contract: top-k and NMS order candidates by (score desc, index asc).
 *
 *   prev      fp32 [K, R, ld_prev]; only the gt_classes columns are read
 *   boxes     fp32 [R, 4]  view-1 proposal boxes
 *   gt_classes int32 [G] sorted ascending, G >= 1
 *   gt_count  NULL, or a DEVICE int32: only the first *gt_count entries of gt_classes are live (G is then
 *             the capacity, e.g. C) -- lets the caller keep the image-level labels on the device
 *             (soswsod_image_level_gt) instead of reading torch.unique's size back to the host
 *   top_k = max(int(R * WSL.MIST_P), 1) computed by the caller with the reference's Python arithmetic
 *           (roi_heads_oicrplus.py:657-662); score_thr/nms_thr = WSL.MIST_THRE (0.05) / 0.01
 *   iou_lo/iou_hi = MODEL.ROI_HEADS.IOU_THRESHOLDS (0.5, 0.6): <lo background, [lo,hi) ignore, >=hi fg
 *   outputs (per branch k):
 *     seed_count int32 [K]; seed_index int32 [K, max_seeds]; seed_class int32 [K, max_seeds];
 *     seed_score fp32 [K, max_seeds]  (score-descending; max_seeds = top_k*G)
 *     gt_class int32 [K, R] (C = background, -1 = ignore); gt_weight fp32 [K, R] (matched seed's score,
 *     NOT yet zeroed for ignored rows); gt_index int32 [K, R] (matched seed's proposal index);
 *     counts int32 [K, 3] = (#fg, #bg, #ignore).
 *   workspace: soswsod_oicr_mine_workspace_bytes(top_k, G, K).
 * ------------------------------------------------------------------------------------------- */
size_t soswsod_oicr_mine_workspace_bytes(int top_k, int G, int K);
int soswsod_oicr_mine_label(const float* prev, long long ld_prev, const float* boxes, const int32_t* gt_classes,
                            int G, const int32_t* gt_count, int R, int C, int K, int top_k, float score_thr, float nms_thr,
                            float iou_lo, float iou_hi, int32_t* seed_count, int32_t* seed_index,
                            int32_t* seed_class, float* seed_score, int32_t* gt_class, float* gt_weight,
                            int32_t* gt_index, int32_t* counts, void* workspace, size_t workspace_bytes,
                            soswsod_stream_t stream);

/* soswsod_oicr_loss: weighted CE + L1 box regression and their gradients for all K branches and V views.
 *   Replaces OICROutputs.softmax_cross_entropy_loss / box_reg_loss / _log_accuracy
 *   (W/modeling/roi_heads/fast_rcnn_oicr.py:157-352), Box2BoxTransform.get_deltas
 *   (D/modeling/box_regression.py:38-71) and the per-branch 4-view mean of roi_heads_oicrplus.py:378-388.
 *
 *   logits   fp32 [V*R, ld]; branch k: class logits in [col_ref0 + k*ref_stride, +C+1), box deltas in the
 *            next 4C columns
 *   boxes    fp32 [V, R, 4]; gt boxes of view v are boxes[v][gt_index]
 *   flip_quirk != 0 reproduces roi_heads_oicrplus.py:381 (the last view's loss uses view 2's predictions)
 *   losses   fp32 [K, 2] = (loss_cls_r{k}, loss_box_reg_r{k}), each the mean over the V views
 *   view_losses fp32 [K, V, 2] (may be NULL)
 *   acc_counts int32 [K, V, 5] = (#instances, #fg, #accurate, #fg accurate, #false negative) (may be NULL)
 *   dlogits  fp32 [V*R, ld_d] or NULL: unit-upstream gradient written into the same column blocks
 *            (every element of the K refinement blocks is written, zeros included).
 * ------------------------------------------------------------------------------------------- */
int soswsod_oicr_loss(const float* logits, long long ld, int col_ref0, int ref_stride, const float* boxes,
                      const int32_t* gt_class, const float* gt_weight, const int32_t* gt_index, int num_views,
                      int R, int C, int K, int flip_quirk, float wx, float wy, float ww, float wh, float* losses,
                      float* view_losses, int32_t* acc_counts, float* dlogits, long long ld_d,
                      soswsod_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (5) Test-time scoring, TTA merge, per-class bitmask NMS.
 *
 * soswsod_predict: predict_probs_K / predict_boxes_K + apply_deltas
 *   (fast_rcnn_oicr.py:674-735, D/modeling/box_regression.py:73-110):
 *   probs [R, C+1] = mean_k softmax(logits_k); pred_boxes [R, 4C] = apply_deltas(mean_k deltas_k, boxes).
 * ------------------------------------------------------------------------------------------- */
int soswsod_predict(const float* logits, long long ld, int col_ref0, int ref_stride, const float* boxes, int R,
                    int C, int K, float wx, float wy, float ww, float wh, float* probs, float* pred_boxes,
                    soswsod_stream_t stream);

/* soswsod_tta_accumulate: inverse view transform (undo h-flip about view_w, then scale by
 * (scale_x, scale_y)) and running sum of boxes/probs over views; `finalize_div` > 0 divides the
 * accumulators by it after adding (pass num_views on the last view).  Replaces the CPU-numpy loop of
 * W/modeling/test_time_augmentation_avg.py:349-371.  first != 0 overwrites instead of accumulating. */
int soswsod_tta_accumulate(const float* pred_boxes, const float* probs, int R, int C, float scale_x,
                           float scale_y, int flipped, float view_w, int first, float finalize_div,
                           float* acc_boxes, float* acc_probs, soswsod_stream_t stream);

/* Test-time augmentation for ALL views of an image in one launch each.
 * A view is described by SOSWSOD_TTA_VIEW_PARAMS floats in a HOST array (copied into the launch
 * parameters, so nothing is read back and the call stays stream-ordered):
 *   [0] fwd_sx = fp32(new_w * 1.0 / w)   [1] fwd_sy = fp32(new_h * 1.0 / h)      ResizeTransform.apply_coords
 *   [2] flipped (0 / 1)                  [3] view width new_w   [4] view height new_h
 *   [5] value written into the roi's batch-index column (index of the view inside its feature tensor)
 *   [6] inv_sx = fp32(w * 1.0 / new_w)   [7] inv_sy = fp32(h * 1.0 / new_h)      ResizeTransform.inverse()
 *   [8] post_sx, [9] post_sy: inverse of the reference's pre-transform (:169-173, stored image -> dataset
 *       size), applied after [6], [7] by soswsod_tta_merge; 1 when the stored image is at the dataset's size
 *
 * soswsod_tta_views: the proposal path of DatasetMapperTTAAVG.__call__ / transform_proposals
 *   (W/modeling/test_time_augmentation_avg.py:29-71, 186-195): TransformList([Resize, HFlip]).apply_box
 *   (corner min / max after every member), Boxes.clip to the view, Boxes.nonempty(min_box_size)
 *   (D/structures/boxes.py:183-210).  boxes fp32 [R,4] in original-image coordinates ->
 *   rois fp32 [V*R, 5] view-major (batch index, x1, y1, x2, y2), keep uint8 [V, R] (the nonempty
 *   mask), dropped int32 [V] = proposals of the view that fail it (the reference would drop them and
 *   then fail to average views of different length: the caller treats a non-zero count as an error).
 *
 * soswsod_tta_merge: GeneralizedRCNNWithTTAAVG._get_augmented_boxes (:349-371): per view
 *   tfm.inverse().apply_box (inverse flip about new_w, then * inv_sx / inv_sy), then the mean over views
 *   (sum in view order, one division by V) of pred_boxes [V, R, 4C] -> [R, 4C] and probs [V, R, C+1] ->
 *   [R, C+1].  Replaces V device->host->device numpy round trips per image. */
#define SOSWSOD_TTA_MAX_VIEWS 32
#define SOSWSOD_TTA_VIEW_PARAMS 10
int soswsod_tta_views(const float* boxes, int R, const float* view_params_host, int V, float min_box_size,
                      float* rois, uint8_t* keep, int32_t* dropped, soswsod_stream_t stream);
int soswsod_tta_merge(const float* pred_boxes, const float* probs, int V, int R, int C,
                      const float* view_params_host, float* mean_boxes, float* mean_probs,
                      soswsod_stream_t stream);

/* soswsod_nms: greedy NMS, suppress j when IoU(i,j) > thr (strict); keep = original indices in score-
 * descending order (ties: lower index first).  Replaces torchvision `nms` (D/layers/nms.py:6-7,25);
 * bitmask algorithm as D/layers/csrc/nms_rotated/nms_rotated_cuda.cu:21-143 but with the reduce on the
 * device.  n <= 16384. */
size_t soswsod_nms_workspace_bytes(int n);
int soswsod_nms(const float* boxes, const float* scores, int n, float iou_thr, int64_t* keep, int32_t* num_keep,
                void* workspace, size_t workspace_bytes, soswsod_stream_t stream);

/* soswsod_detect: fast_rcnn_inference_single_image (fast_rcnn_oicr.py:86-148) with per-class NMS
 * (D/layers/nms.py:10-29, per-class variant): finite-row filter, drop background column, clip to
 * (img_h, img_w), score > score_thr, per-class NMS(nms_thr), global top-`topk` by score.
 *   probs [R, C+1], pred_boxes [R, 4C];
 *   det_boxes fp32 [topk, 4]; det_scores fp32 [topk]; det_classes int32 [topk]; det_rows int32 [topk];
 *   num_det int32 [1].  R <= 16384. */
size_t soswsod_detect_workspace_bytes(int R, int C);
int soswsod_detect(const float* probs, const float* pred_boxes, int R, int C, float img_h, float img_w,
                   float score_thr, float nms_thr, int topk, float* det_boxes, float* det_scores,
                   int32_t* det_classes, int32_t* det_rows, int32_t* num_det, void* workspace,
                   size_t workspace_bytes, soswsod_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (next) SGD + momentum + weight decay with the per-group lr / wd of
 * uwsod/detectron2/solver/build.py:143-218 chosen by the caller:
 *   g = grad*grad_scale + weight_decay*p ; buf = momentum*buf + g ; p -= lr*buf
 * and, if param_bf16 != NULL, the refreshed bf16 GEMM-operand copy of p in the same pass. Contiguous fp32. */
int soswsod_sgd_step(float* param, const float* grad, float* momentum_buf, long long n, float lr, float momentum,
                     float weight_decay, float grad_scale, void* param_bf16, soswsod_stream_t stream);

/* The same update for up to SOSWSOD_SGD_MAX_TENSORS tensors (or shards: pass pre-offset pointers) in ONE launch --
 * the whole `optimizer.step()` of uwsod/projects/WSL/tools/train_net_multi.py:157-164 for the head's 8 + 4K
 * parameters.  `tensors` is a HOST array (copied into the launch parameters).  Per tensor: lr / weight_decay of its
 * parameter group; out_bf16 (may be NULL) receives the refreshed bf16 GEMM operand, out_f32 (may be NULL) an fp32
 * copy of the new value (the fused head-bias vector). */
#define SOSWSOD_SGD_MAX_TENSORS 32
typedef struct soswsod_sgd_tensor {
    float* param;
    const float* grad;
    float* momentum_buf;
    void* out_bf16;
    float* out_f32;
    long long n;
    float lr;
    float weight_decay;
} soswsod_sgd_tensor;
int soswsod_sgd_multi(const soswsod_sgd_tensor* tensors, int count, float momentum, float grad_scale,
                      soswsod_stream_t stream);

/* The data-parallel form of the same step for the big weight matrices, fused with its collectives over NVSwitch
 * multicast (NVLS) memory -- replaces DistributedDataParallel's gradient all-reduce (tools/train_net_multi.py:75-78) AND
 * optimizer.step() (:157-164) for those tensors with one kernel per rank: for the rows a rank owns,
 *   grad = grad_scale * (sum over ranks, read with multimem.ld_reduce from the multicast address grad_mc),
 *   SGD + momentum + weight decay on the local fp32 rows, and the refreshed bf16 operand rows stored with multimem.st
 *   to the multicast address out_bf16_mc (they land in every rank's operand matrix).
 * param / momentum_buf: local device pointers; grad_mc / out_bf16_mc: multicast virtual addresses of symmetric buffers
 * (cuMulticast* / torch symmetric memory), all pre-offset to the owned rows; n % 4 == 0, fp32 pointers 16-byte aligned.  The caller
 * provides the cross-rank barriers: every rank's gradient complete before the launch; all launches complete before
 * any rank reads the operands or overwrites the gradients.  `tensors` is a HOST array. */
#define SOSWSOD_SGD_NVLS_MAX_TENSORS 8
typedef struct soswsod_sgd_nvls_tensor {
    float* param;
    const float* grad_mc;
    float* momentum_buf;
    void* out_bf16_mc;
    long long n;
    float lr;
    float weight_decay;
} soswsod_sgd_nvls_tensor;
int soswsod_sgd_nvls(const soswsod_sgd_nvls_tensor* tensors, int count, float momentum, float grad_scale,
                     int max_ctas /* persistent grid size; <= 0: 1 per SM */, soswsod_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (6) PGF, the consumer of the detection-results json.  Replaces the per-image
 *     Python loops of tools/pgf.py:221-270 (`pgf`) with `contain_cal` (:210-219), in the same double
 *     precision: detections of image i are rows [img_offsets[i], img_offsets[i+1]) in list order;
 *     keep[r] = 1 when row r survives both the keep-threshold rule (first detection of a category
 *     always kept, later ones need score >= t_keep) and the same-category containment rule
 *     (area(r ∩ j) / (area(r) + 1e-6) >= t_con drops r; categories in the 128-bit diff mask are exempt
 *     unless use_diff).  boxes are XYWH doubles [n,4], 32-byte aligned.
 * ------------------------------------------------------------------------------------------- */
int soswsod_pgf(const double* boxes_xywh, const double* scores, const int* categories, const int* img_offsets,
                int num_images, double t_con, double t_keep, int use_diff, unsigned long long diff_mask_lo,
                unsigned long long diff_mask_hi, unsigned char* keep, soswsod_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SOSWSOD_B200_H_ */
