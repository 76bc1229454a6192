/*
 * Scalar C restatement of the third-party operators the reference's hot path leans on
 * (un-vendored torchvision: roi_pool and nms).  TEST INFRASTRUCTURE ONLY -- nothing under
 * sos_wsod_b200/ links or loads this file; it is compiled by oracle/Makefile into
 * oracle/_build/libref_kernels.so and loaded with ctypes from tests/ (and bench.py's CPU arm).
 *
 * Algorithm sources (published semantics):
 *   - roi_pool: torchvision `roi_pool` (pinned 0.7/0.10 by the reference, call site
 *     uwsod/projects/WSL/wsl/modeling/poolers.py:183-186,263-270); the same bin arithmetic is restated
 *     in-repo by uwsod/projects/WSL/wsl/layers/csrc/ROILoopPool/ROILoopPool_cuda.cu:77-137.
 *   - nms: torchvision `nms` (call site uwsod/detectron2/layers/nms.py:6-7,20,25); sequential semantics
 *     as in uwsod/detectron2/layers/csrc/nms_rotated/nms_rotated_cpu.cpp:8-60 but axis-aligned IoU and
 *     strict `>` (torchvision), not `>=`.
 *   - pairwise IoU: uwsod/detectron2/structures/boxes.py:329-361.
 * Cross-checked bit-for-bit against torchvision 0.26's CPU kernels in tests/test_oracle.py.
 *
 * Build flags matter: -O2 -ffp-contract=off (no FMA contraction; fp32 results must be IEEE-exact).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* feat [N,C,H,W] fp32, rois [R,5] = (batch, x1, y1, x2, y2); out [R,C,P,P]; argmax [R,C,P,P] int32
 * (index h*W+w inside the (n,c) plane, -1 for an empty bin). */
void ref_roi_pool_forward(const float* feat, int N, int C, int H, int W, const float* rois, int R,
                          int PH, int PW, float spatial_scale, float* out, int32_t* argmax) {
    (void)N;
    for (int r = 0; r < R; ++r) {
        const float* roi = rois + 5 * r;
        int b = (int)roi[0];
        int rs_w = (int)roundf(roi[1] * spatial_scale);
        int rs_h = (int)roundf(roi[2] * spatial_scale);
        int re_w = (int)roundf(roi[3] * spatial_scale);
        int re_h = (int)roundf(roi[4] * spatial_scale);
        int roi_w = imax(re_w - rs_w + 1, 1);
        int roi_h = imax(re_h - rs_h + 1, 1);
        float bin_h = (float)roi_h / (float)PH;
        float bin_w = (float)roi_w / (float)PW;
        for (int c = 0; c < C; ++c) {
            const float* plane = feat + ((size_t)b * C + c) * H * W;
            for (int ph = 0; ph < PH; ++ph) {
                int hs = (int)floorf((float)ph * bin_h);
                int he = (int)ceilf((float)(ph + 1) * bin_h);
                hs = imin(imax(hs + rs_h, 0), H);
                he = imin(imax(he + rs_h, 0), H);
                for (int pw = 0; pw < PW; ++pw) {
                    int ws = (int)floorf((float)pw * bin_w);
                    int we = (int)ceilf((float)(pw + 1) * bin_w);
                    ws = imin(imax(ws + rs_w, 0), W);
                    we = imin(imax(we + rs_w, 0), W);
                    int empty = (he <= hs) || (we <= ws);
                    float maxval = empty ? 0.f : -FLT_MAX;
                    int maxidx = -1;
                    for (int h = hs; h < he; ++h)
                        for (int w = ws; w < we; ++w) {
                            int idx = h * W + w;
                            if (plane[idx] > maxval) { maxval = plane[idx]; maxidx = idx; }
                        }
                    size_t o = (((size_t)r * C + c) * PH + ph) * PW + pw;
                    out[o] = maxval;
                    argmax[o] = maxidx;
                }
            }
        }
    }
}

/* grad_feat [N,C,H,W] must be zeroed by the caller. */
void ref_roi_pool_backward(const float* grad_out, const int32_t* argmax, const float* rois, int R, int C,
                           int H, int W, int PH, int PW, float* grad_feat) {
    for (int r = 0; r < R; ++r) {
        int b = (int)rois[5 * r];
        for (int c = 0; c < C; ++c) {
            float* plane = grad_feat + ((size_t)b * C + c) * H * W;
            size_t o = ((size_t)r * C + c) * PH * PW;
            for (int i = 0; i < PH * PW; ++i) {
                int a = argmax[o + i];
                if (a != -1) plane[a] += grad_out[o + i];
            }
        }
    }
}

static inline float box_iou(const float* a, const float* b) {
    float area_a = (a[2] - a[0]) * (a[3] - a[1]);
    float area_b = (b[2] - b[0]) * (b[3] - b[1]);
    float w = fminf(a[2], b[2]) - fmaxf(a[0], b[0]);
    float h = fminf(a[3], b[3]) - fmaxf(a[1], b[1]);
    if (w < 0.f) w = 0.f;
    if (h < 0.f) h = 0.f;
    float inter = w * h;
    return inter / (area_a + area_b - inter);
}

typedef struct { float s; int i; } si_t;
static int cmp_desc_stable(const void* pa, const void* pb) {
    const si_t* a = (const si_t*)pa; const si_t* b = (const si_t*)pb;
    if (a->s > b->s) return -1;
    if (a->s < b->s) return 1;
    return a->i - b->i;
}

/* Greedy NMS: boxes [n,4] XYXY, scores [n]; keep_out receives original indices in score-descending
 * order; returns the number kept.  Suppress j when IoU(i,j) > thr (strict). */
int ref_nms(const float* boxes, const float* scores, int n, float thr, int64_t* keep_out) {
    si_t* order = (si_t*)malloc(sizeof(si_t) * (size_t)(n > 0 ? n : 1));
    unsigned char* dead = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
    for (int i = 0; i < n; ++i) { order[i].s = scores[i]; order[i].i = i; }
    qsort(order, (size_t)n, sizeof(si_t), cmp_desc_stable);
    int nk = 0;
    for (int a = 0; a < n; ++a) {
        if (dead[a]) continue;
        int i = order[a].i;
        keep_out[nk++] = i;
        for (int b = a + 1; b < n; ++b) {
            if (dead[b]) continue;
            if (box_iou(boxes + 4 * i, boxes + 4 * order[b].i) > thr) dead[b] = 1;
        }
    }
    free(order); free(dead);
    return nk;
}

/* pairwise IoU of a [M,4] vs b [R,4] -> out [M,R]; 0 where the intersection is empty. */
void ref_pairwise_iou(const float* a, int M, const float* b, int R, float* out) {
    for (int m = 0; m < M; ++m)
        for (int r = 0; r < R; ++r) {
            const float* p = a + 4 * m; const float* q = b + 4 * r;
            float w = fminf(p[2], q[2]) - fmaxf(p[0], q[0]);
            float h = fminf(p[3], q[3]) - fmaxf(p[1], q[1]);
            if (w < 0.f) w = 0.f;
            if (h < 0.f) h = 0.f;
            float inter = w * h;
            float area_a = (p[2] - p[0]) * (p[3] - p[1]);
            float area_b = (q[2] - q[0]) * (q[3] - q[1]);
            out[(size_t)m * R + r] = inter > 0.f ? inter / (area_a + area_b - inter) : 0.f;
        }
}

/* ------------------------------------------------------------------------------------------------
 * Test-time augmentation: scalar restatement of
 *   - transform_proposals (uwsod/projects/WSL/wsl/modeling/test_time_augmentation_avg.py:29-71):
 *     TransformList([ResizeTransform, HFlipTransform]).apply_box on fp32 boxes -- ResizeTransform.apply_coords
 *     (uwsod/detectron2/data/transforms/transform.py:123-126) multiplies by the double new_w / w rounded to fp32;
 *     fvcore Transform.apply_box (un-vendored, published) takes the min / max of the 4 transformed corners after
 *     every member; HFlipTransform: x -> width - x -- then Boxes.clip and Boxes.nonempty
 *     (uwsod/detectron2/structures/boxes.py:183-210);
 *   - the per-view inverse of GeneralizedRCNNWithTTAAVG._get_augmented_boxes (:353-365).
 * ------------------------------------------------------------------------------------------------ */
static inline float fminf2(float a, float b) { return a < b ? a : b; }
static inline float fmaxf2(float a, float b) { return a > b ? a : b; }
static inline float clampf(float v, float lo, float hi) { return fminf2(fmaxf2(v, lo), hi); }

/* boxes [n,4] in the stored image (h x w) -> out [n,4] in the view (new_h x new_w), keep [n] */
void ref_tta_transform_proposals(const float* boxes, int n, int h, int w, int new_h, int new_w, int flipped,
                                 float min_box_size, float* out, uint8_t* keep) {
    const float sx = (float)((double)new_w * 1.0 / (double)w);
    const float sy = (float)((double)new_h * 1.0 / (double)h);
    for (int i = 0; i < n; ++i) {
        const float* b = boxes + 4 * i;
        float xa = b[0] * sx, xb = b[2] * sx, ya = b[1] * sy, yb = b[3] * sy;
        float x1 = fminf2(xa, xb), x2 = fmaxf2(xa, xb), y1 = fminf2(ya, yb), y2 = fmaxf2(ya, yb);
        if (flipped) {
            xa = (float)new_w - x1;
            xb = (float)new_w - x2;
            x1 = fminf2(xa, xb);
            x2 = fmaxf2(xa, xb);
        }
        x1 = clampf(x1, 0.f, (float)new_w);
        x2 = clampf(x2, 0.f, (float)new_w);
        y1 = clampf(y1, 0.f, (float)new_h);
        y2 = clampf(y2, 0.f, (float)new_h);
        out[4 * i + 0] = x1;
        out[4 * i + 1] = y1;
        out[4 * i + 2] = x2;
        out[4 * i + 3] = y2;
        keep[i] = ((x2 - x1) > min_box_size) && ((y2 - y1) > min_box_size);
    }
}

/* boxes [n,4] in the view -> out [n,4] in the stored image; post_w / post_h > 0: then on to the dataset's size
 * (the reference's pre-transform, :169-173), else pass 0 */
void ref_tta_inverse_boxes(const float* boxes, int n, int h, int w, int new_h, int new_w, int flipped, int post_h,
                           int post_w, float* out) {
    const float sx = (float)((double)w * 1.0 / (double)new_w);
    const float sy = (float)((double)h * 1.0 / (double)new_h);
    for (int i = 0; i < n; ++i) {
        const float* b = boxes + 4 * i;
        float x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
        if (flipped) {
            const float xa = (float)new_w - x1, xb = (float)new_w - x2;
            x1 = fminf2(xa, xb);
            x2 = fmaxf2(xa, xb);
        }
        float xa = x1 * sx, xb = x2 * sx, ya = y1 * sy, yb = y2 * sy;
        x1 = fminf2(xa, xb);
        x2 = fmaxf2(xa, xb);
        y1 = fminf2(ya, yb);
        y2 = fmaxf2(ya, yb);
        if (post_w > 0 && post_h > 0) {
            const float px = (float)((double)post_w * 1.0 / (double)w), py = (float)((double)post_h * 1.0 / (double)h);
            xa = x1 * px;
            xb = x2 * px;
            ya = y1 * py;
            yb = y2 * py;
            x1 = fminf2(xa, xb);
            x2 = fmaxf2(xa, xb);
            y1 = fminf2(ya, yb);
            y2 = fmaxf2(ya, yb);
        }
        out[4 * i + 0] = x1;
        out[4 * i + 1] = y1;
        out[4 * i + 2] = x2;
        out[4 * i + 3] = y2;
    }
}
