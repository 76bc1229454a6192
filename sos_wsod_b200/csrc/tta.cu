// Test-time augmentation on the device:
//   tta_views   -- DatasetMapperTTAAVG's proposal path (W/wsl/modeling/test_time_augmentation_avg.py:29-71,186-195):
//                  for every view, ResizeTransform.apply_box (U/detectron2/data/transforms/transform.py:123-126:
//                  x * fp32(new_w / w), y * fp32(new_h / h)) then HFlipTransform.apply_box (x -> new_w - x), each
//                  followed by the corner min / max of Transform.apply_box, Boxes.clip to the view
//                  (U/detectron2/structures/boxes.py:183-196) and Boxes.nonempty(threshold)
//                  (boxes.py:198-210: both sides > threshold).  Writes the [V*R, 5] roi list the pooler takes,
//                  a keep mask and the number of proposals a view would drop -- the reference drops them, which
//                  breaks its own mean over views (all views must keep the same rows), so the caller checks the
//                  counts instead of compacting.
//   tta_merge   -- GeneralizedRCNNWithTTAAVG._get_augmented_boxes (:349-371) for ALL views in one launch: inverse
//                  transform of every predicted box (inverse flip about the view width, corner min / max, inverse
//                  resize x * fp32(w / new_w)), then the mean over views of boxes and class probabilities
//                  (sum in view order, divided by V).
//                  A view whose source image was itself resized from the dataset's size (the reference's pre_tfm,
//                  :169-173) ends its inverse with a second multiply (params 8, 9; 1 otherwise).
// Both replace per-view CPU-numpy round trips (:57, :358) by one launch per image.
#include "common.cuh"

namespace soswsod {

struct TtaViews {
    int V;
    float p[SOSWSOD_TTA_MAX_VIEWS][SOSWSOD_TTA_VIEW_PARAMS];
};

__global__ void tta_views_kernel(const float* __restrict__ boxes, int R, const __grid_constant__ TtaViews vs,
                                 float min_box_size, float* __restrict__ rois, uint8_t* __restrict__ keep,
                                 int32_t* __restrict__ dropped) {
    const int v = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const float sx = vs.p[v][0], sy = vs.p[v][1];
    const bool flip = vs.p[v][2] != 0.f;
    const float vw = vs.p[v][3], vh = vs.p[v][4], bidx = vs.p[v][5];
    bool drop = false;
    if (r < R) {
        const float4 b = reinterpret_cast<const float4*>(boxes)[r];
        // resize: the four corners scaled, then min / max
        float xa = __fmul_rn(b.x, sx), xb = __fmul_rn(b.z, sx);
        float ya = __fmul_rn(b.y, sy), yb = __fmul_rn(b.w, sy);
        float x1 = fminf(xa, xb), x2 = fmaxf(xa, xb);
        float y1 = fminf(ya, yb), y2 = fmaxf(ya, yb);
        if (flip) {
            xa = __fsub_rn(vw, x1);
            xb = __fsub_rn(vw, x2);
            x1 = fminf(xa, xb);
            x2 = fmaxf(xa, xb);
        }
        x1 = fminf(fmaxf(x1, 0.f), vw);
        x2 = fminf(fmaxf(x2, 0.f), vw);
        y1 = fminf(fmaxf(y1, 0.f), vh);
        y2 = fminf(fmaxf(y2, 0.f), vh);
        const bool ok = (__fsub_rn(x2, x1) > min_box_size) && (__fsub_rn(y2, y1) > min_box_size);
        drop = !ok;
        float* o = rois + ((size_t)v * R + r) * 5;
        o[0] = bidx;
        o[1] = x1;
        o[2] = y1;
        o[3] = x2;
        o[4] = y2;
        keep[(size_t)v * R + r] = ok ? 1 : 0;
    }
    const unsigned bal = __ballot_sync(FULL_MASK, drop);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(dropped + v, __popc(bal));   // integer bookkeeping
}

__global__ void tta_merge_kernel(const float* __restrict__ pred_boxes, const float* __restrict__ probs, long long nbox,
                                 long long nprob, const __grid_constant__ TtaViews vs, float* __restrict__ mean_boxes,
                                 float* __restrict__ mean_probs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int V = vs.V;
    const float fV = (float)V;
    if (i < nbox) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int v = 0; v < V; ++v) {
            const float4 b = reinterpret_cast<const float4*>(pred_boxes)[(size_t)v * nbox + i];
            float x1 = b.x, x2 = b.z;
            if (vs.p[v][2] != 0.f) {
                const float xa = __fsub_rn(vs.p[v][3], b.x), xb = __fsub_rn(vs.p[v][3], b.z);
                x1 = fminf(xa, xb);
                x2 = fmaxf(xa, xb);
            }
            float xa = __fmul_rn(x1, vs.p[v][6]), xb = __fmul_rn(x2, vs.p[v][6]);
            float ya = __fmul_rn(b.y, vs.p[v][7]), yb = __fmul_rn(b.w, vs.p[v][7]);
            x1 = fminf(xa, xb);
            x2 = fmaxf(xa, xb);
            float y1 = fminf(ya, yb), y2 = fmaxf(ya, yb);
            // the reference's pre-transform inverse (stored image -> dataset size); factors are 1 when there is none
            xa = __fmul_rn(x1, vs.p[v][8]);
            xb = __fmul_rn(x2, vs.p[v][8]);
            ya = __fmul_rn(y1, vs.p[v][9]);
            yb = __fmul_rn(y2, vs.p[v][9]);
            a.x = __fadd_rn(a.x, fminf(xa, xb));
            a.y = __fadd_rn(a.y, fminf(ya, yb));
            a.z = __fadd_rn(a.z, fmaxf(xa, xb));
            a.w = __fadd_rn(a.w, fmaxf(ya, yb));
        }
        a.x = __fdiv_rn(a.x, fV);
        a.y = __fdiv_rn(a.y, fV);
        a.z = __fdiv_rn(a.z, fV);
        a.w = __fdiv_rn(a.w, fV);
        reinterpret_cast<float4*>(mean_boxes)[i] = a;
    }
    if (i < nprob) {
        float a = 0.f;
        for (int v = 0; v < V; ++v) a = __fadd_rn(a, probs[(size_t)v * nprob + i]);
        mean_probs[i] = __fdiv_rn(a, fV);
    }
}

static int fill_views(const float* view_params, int V, TtaViews* out) {
    out->V = V;
    for (int v = 0; v < V; ++v)
        for (int k = 0; k < SOSWSOD_TTA_VIEW_PARAMS; ++k) out->p[v][k] = view_params[v * SOSWSOD_TTA_VIEW_PARAMS + k];
    return 0;
}

}  // namespace soswsod

using namespace soswsod;

extern "C" int soswsod_tta_views(const float* boxes, int R, const float* view_params_host, int V, float min_box_size,
                                 float* rois, uint8_t* keep, int32_t* dropped, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(boxes && view_params_host && rois && keep && dropped, "tta_views: null pointer");
    SOSWSOD_CHECK_ARG(R > 0 && V > 0 && V <= SOSWSOD_TTA_MAX_VIEWS, "tta_views: R > 0 and 1 <= V <= %d required", SOSWSOD_TTA_MAX_VIEWS);
    SOSWSOD_CHECK_ARG(((uintptr_t)boxes & 15) == 0, "tta_views: boxes must be 16B aligned");
    TtaViews vs;
    fill_views(view_params_host, V, &vs);
    SOSWSOD_CHECK_CUDA(cudaMemsetAsync(dropped, 0, sizeof(int32_t) * V, (cudaStream_t)stream));
    tta_views_kernel<<<dim3((R + 255) / 256, V), 256, 0, (cudaStream_t)stream>>>(boxes, R, vs, min_box_size, rois, keep, dropped);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

extern "C" int soswsod_tta_merge(const float* pred_boxes, const float* probs, int V, int R, int C,
                                 const float* view_params_host, float* mean_boxes, float* mean_probs,
                                 soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(pred_boxes && probs && view_params_host && mean_boxes && mean_probs, "tta_merge: null pointer");
    SOSWSOD_CHECK_ARG(R > 0 && C > 0 && V > 0 && V <= SOSWSOD_TTA_MAX_VIEWS, "tta_merge: bad shape (1 <= V <= %d)", SOSWSOD_TTA_MAX_VIEWS);
    SOSWSOD_CHECK_ARG(((uintptr_t)pred_boxes & 15) == 0 && ((uintptr_t)mean_boxes & 15) == 0, "tta_merge: boxes must be 16B aligned");
    TtaViews vs;
    fill_views(view_params_host, V, &vs);
    const long long nbox = (long long)R * C, nprob = (long long)R * (C + 1);
    const long long total = nbox > nprob ? nbox : nprob;
    tta_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pred_boxes, probs, nbox, nprob, vs,
                                                                                      mean_boxes, mean_probs);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}
