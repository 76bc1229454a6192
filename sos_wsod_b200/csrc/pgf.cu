// PGF (pseudo-ground-truth filter) of the detection results on the device: the step that consumes the json the hot
// path writes.  Restates tools/pgf.py:221-270 (`pgf`) and :210-219 (`contain_cal`) of the
// reference, which run as an O(n^2) Python loop per image: one CTA per image, double-precision arithmetic exactly as
// Python evaluates it (built with -fmad=false; every operation is a separate IEEE double op).
//
//   stage 1  keep rule: walking an image's detections in list order, the first detection of every category is always
//            kept; any later one is dropped when score < t_keep.
//   stage 2  containment: a survivor i is dropped when another survivor j of the same category has
//            area(i ∩ j) / (area(i) + 1e-6) >= t_con   (boxes are XYWH, as pgf.py treats the json's `bbox`);
//            categories listed in diff_mask are exempt unless use_diff.
#include "common.cuh"

namespace soswsod {

constexpr int kPgfThreads = 128;

__device__ __forceinline__ double pgf_contain(const double4 a, const double4 b) {
    const double a2 = __dadd_rn(a.z, a.x), a3 = __dadd_rn(a.w, a.y);
    const double b2 = __dadd_rn(b.z, b.x), b3 = __dadd_rn(b.w, b.y);
    const double c0 = fmax(a.x, b.x), c1 = fmax(a.y, b.y), c2 = fmin(a2, b2), c3 = fmin(a3, b3);
    const double dw = __dsub_rn(c2, c0), dh = __dsub_rn(c3, c1);
    const double area_c = __dmul_rn(dw > 0.0 ? dw : 0.0, dh > 0.0 ? dh : 0.0);
    const double aw = __dsub_rn(a2, a.x), ah = __dsub_rn(a3, a.y);
    const double area_a = __dmul_rn(aw > 0.0 ? aw : 0.0, ah > 0.0 ? ah : 0.0);
    return __ddiv_rn(area_c, __dadd_rn(area_a, 1e-6));
}

__global__ void __launch_bounds__(kPgfThreads)
pgf_kernel(const double* __restrict__ boxes, const double* __restrict__ scores, const int* __restrict__ cats,
           const int* __restrict__ img_offsets, double t_con, double t_keep, int use_diff,
           unsigned long long diff_lo, unsigned long long diff_hi, unsigned char* __restrict__ keep) {
    const int img = blockIdx.x;
    const int lo = img_offsets[img], hi = img_offsets[img + 1];
    // stage 1: first of its category in list order, or score >= t_keep
    for (int i = lo + threadIdx.x; i < hi; i += kPgfThreads) {
        const int c = cats[i];
        bool first = true;
        for (int j = lo; j < i; ++j)
            if (cats[j] == c) {
                first = false;
                break;
            }
        keep[i] = (first || !(scores[i] < t_keep)) ? 1 : 0;
    }
    __syncthreads();
    // stage 2: containment among the survivors (decisions use the stage-1 list, never each other)
    for (int i = lo + threadIdx.x; i < hi; i += kPgfThreads) {
        if (!keep[i]) continue;
        const int c = cats[i];
        const bool is_diff = c >= 0 && c < 128 && (((c < 64 ? diff_lo >> c : diff_hi >> (c - 64)) & 1ull) != 0);
        if (!use_diff && is_diff) continue;
        const double4 a = *reinterpret_cast<const double4*>(boxes + 4 * (size_t)i);
        bool save = true;
        for (int j = lo; j < hi && save; ++j) {
            if (j == i || cats[j] != c || keep[j] == 0) continue;
            const double4 b = *reinterpret_cast<const double4*>(boxes + 4 * (size_t)j);
            if (pgf_contain(a, b) >= t_con) save = false;
        }
        if (!save) keep[i] = 2;   // dropped in stage 2: still a survivor of stage 1 for everybody else's test
    }
    __syncthreads();
    for (int i = lo + threadIdx.x; i < hi; i += kPgfThreads) keep[i] = keep[i] == 1 ? 1 : 0;
}

}  // namespace soswsod

using namespace soswsod;

extern "C" int soswsod_pgf(const double* boxes_xywh, const double* scores, const int* categories,
                           const int* img_offsets, int num_images, double t_con, double t_keep, int use_diff,
                           unsigned long long diff_mask_lo, unsigned long long diff_mask_hi, unsigned char* keep,
                           soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(num_images >= 0, "pgf: bad image count");
    if (num_images == 0) return SOSWSOD_OK;
    SOSWSOD_CHECK_ARG(boxes_xywh && scores && categories && img_offsets && keep, "pgf: null pointer");
    SOSWSOD_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes_xywh) & 31) == 0, "pgf: boxes must be 32-byte aligned");
    pgf_kernel<<<num_images, kPgfThreads, 0, (cudaStream_t)stream>>>(boxes_xywh, scores, categories, img_offsets, t_con,
                                                                   t_keep, use_diff, diff_mask_lo, diff_mask_hi, keep);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}
