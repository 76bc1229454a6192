// ROI-pool geometry shared by the forward / backward kernels, and the per-call PLAN the fast 7x7 kernels read.
//
// Bin arithmetic follows torchvision roi_pool exactly (restated in-repo by the reference at
// uwsod/projects/WSL/wsl/layers/csrc/ROILoopPool/ROILoopPool_cuda.cu:77-137): C round() of the fp32
// product, fp32 bin size, floor/ceil, clamp to the plane.
//
// Plan layout (device memory owned by the caller, soswsod_roi_pool_plan_bytes()):
//   [0, 512)                      int32 img_start[n + 1]   first position in `order` of every image's rois
//   [512, 512 + 4R)               int32 order[R]           roi indices grouped by image (stable)
//   [align128(512 + 4R), +64R)    RoiRecord[R]             indexed by the ORIGINAL roi index
#pragma once
#include "common.cuh"

namespace soswsod {

struct RoiGeom {
    int batch, rs_w, rs_h;
    float bin_w, bin_h;
};

__device__ __forceinline__ RoiGeom roi_geometry(const float* __restrict__ roi, float scale, int PH, int PW) {
    RoiGeom g;
    g.batch = (int)roi[0];
    g.rs_w = (int)roundf(__fmul_rn(roi[1], scale));
    g.rs_h = (int)roundf(__fmul_rn(roi[2], scale));
    const int re_w = (int)roundf(__fmul_rn(roi[3], scale));
    const int re_h = (int)roundf(__fmul_rn(roi[4], scale));
    const int roi_w = max(re_w - g.rs_w + 1, 1);
    const int roi_h = max(re_h - g.rs_h + 1, 1);
    g.bin_h = __fdiv_rn((float)roi_h, (float)PH);
    g.bin_w = __fdiv_rn((float)roi_w, (float)PW);
    return g;
}

__device__ __forceinline__ int bin_start(int p, float bin, int rs, int limit) {
    return min(max((int)floorf(__fmul_rn((float)p, bin)) + rs, 0), limit);
}
__device__ __forceinline__ int bin_end(int p, float bin, int rs, int limit) {
    return min(max((int)ceilf(__fmul_rn((float)(p + 1), bin)) + rs, 0), limit);
}

// Smallest m >= 1 such that bin p and bin p + m (and anything further apart) never share a cell along one axis.
__device__ __forceinline__ int bin_disjoint_stride(float bin, int rs, int P, int limit) {
    int m = 1;
    for (int p = 0; p + m < P; ++p) {
        const int e = bin_end(p, bin, rs, limit);
        int q = p + m;
        while (q < P && bin_start(q, bin, rs, limit) < e) ++q;
        m = q - p;
    }
    return m;
}

constexpr int kPlanP = 7;             // the plan (and the kernels reading it) is for 7 x 7 bins
constexpr int kPlanMaxImages = 64;
constexpr int kPlanHeaderBytes = 512;

struct __align__(16) RoiRecord {       // 64 bytes
    uint16_t hb[kPlanP][2];            // [ph] = (hstart, hend), clamped to [0, H]
    uint16_t wb[kPlanP][2];            // [pw] = (wstart, wend), clamped to [0, W]
    uint8_t mh, mw;                    // colour strides of the backward (bins mh rows / mw columns apart are disjoint)
    uint16_t batch;
    float scale;                       // row_scale[r] + row_scale_bias (1 without row_scale)
};
static_assert(sizeof(RoiRecord) == 64, "RoiRecord must be 64 bytes");

__host__ __device__ inline size_t plan_records_offset(int R) {
    return ((size_t)kPlanHeaderBytes + 4 * (size_t)R + 127) / 128 * 128;
}
__host__ __device__ inline size_t plan_total_bytes(int R) { return plan_records_offset(R) + 64 * (size_t)R; }

struct PlanView {
    const int* img_start;
    const int* order;
    const RoiRecord* rec;
};
__host__ __device__ inline PlanView plan_view(const void* plan, int R) {
    const uint8_t* p = static_cast<const uint8_t*>(plan);
    PlanView v;
    v.img_start = reinterpret_cast<const int*>(p);
    v.order = reinterpret_cast<const int*>(p + kPlanHeaderBytes);
    v.rec = reinterpret_cast<const RoiRecord*>(p + plan_records_offset(R));
    return v;
}

int device_num_sms();
int device_max_smem();

// roi_pool_fast.cu
int launch_roi_plan(const float* rois, int R, int n, int h, int w, float scale, const float* row_scale, float bias,
                    void* plan, cudaStream_t st);
// returns 1 when the fast kernel was launched, 0 when the shape is outside its envelope (caller falls back to the
// general kernels), negative on error
int launch_fwd_fast(const float* feat, int n, int c, int h, int w, int R, const void* plan, uint16_t* argmax_u16,
                    __nv_bfloat16* out_bf16, long long ld_bf16, cudaStream_t st);
int launch_bwd_fast(const void* grad, int grad_dtype, long long ld_grad, const uint16_t* argmax, int R, const void* plan,
                    int n, int c, int h, int w, float* grad_feat, cudaStream_t st);

}  // namespace soswsod
