// Shared helpers for the sm_100a kernels behind the C ABI in include/soswsod_b200.h.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/soswsod_b200.h"

namespace soswsod {

void set_error(const char* fmt, ...);

#define SOSWSOD_CHECK_ARG(cond, ...)                 \
    do {                                             \
        if (!(cond)) {                               \
            ::soswsod::set_error(__VA_ARGS__);       \
            return SOSWSOD_ERR_INVALID;              \
        }                                            \
    } while (0)

#define SOSWSOD_CHECK_CUDA(expr)                                                              \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::soswsod::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                                 __FILE__, __LINE__);                                         \
            return SOSWSOD_ERR_CUDA;                                                          \
        }                                                                                     \
    } while (0)

#define SOSWSOD_CHECK_LAUNCH() SOSWSOD_CHECK_CUDA(cudaGetLastError())

// cudaFuncAttributeMaxDynamicSharedMemorySize is sticky per (function, device): raise it only when a launch needs more
// than any launch before it (`high_water` = one static int per call site / template instantiation), not on every launch.
#define SOSWSOD_ENSURE_SMEM(kern, bytes)                                                                              \
    do {                                                                                                              \
        static int high_water_[16] = {0};                                                                             \
        int dev_ = 0;                                                                                                 \
        SOSWSOD_CHECK_CUDA(cudaGetDevice(&dev_));                                                                     \
        if (dev_ < 0 || dev_ >= 16 || (int)(bytes) > high_water_[dev_]) {                                             \
            SOSWSOD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            if (dev_ >= 0 && dev_ < 16) high_water_[dev_] = (int)(bytes);                                             \
        }                                                                                                             \
    } while (0)

constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// Block-wide reductions for blocks of up to 1024 threads; `scratch` needs 32 floats. All threads get
// the result. The trailing __syncthreads makes `scratch` reusable immediately.
__device__ __forceinline__ float block_max(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = (lane < nw) ? scratch[lane] : -FLT_MAX;
    r = warp_max(r);
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = (lane < nw) ? scratch[lane] : 0.f;
    r = warp_sum(r);
    __syncthreads();
    return r;
}

// IEEE-exact axis-aligned IoU pieces (no FMA contraction; files using these are built with -fmad=false
// and the intrinsics make the rounding explicit anyway).
__device__ __forceinline__ float box_area_rn(float x1, float y1, float x2, float y2) {
    return __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
}
__device__ __forceinline__ float box_inter_rn(const float4 a, const float4 b) {
    float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    w = fmaxf(w, 0.f);
    h = fmaxf(h, 0.f);
    return __fmul_rn(w, h);
}
// torchvision nms: inter / (a + b - inter), no zero guard.
__device__ __forceinline__ float box_iou_nms_rn(const float4 a, const float4 b) {
    const float inter = box_inter_rn(a, b);
    const float ua = __fsub_rn(__fadd_rn(box_area_rn(a.x, a.y, a.z, a.w), box_area_rn(b.x, b.y, b.z, b.w)), inter);
    return __fdiv_rn(inter, ua);
}
// detectron2 pairwise_iou: 0 where inter == 0.
__device__ __forceinline__ float box_iou_pairwise_rn(const float4 a, float area_a, const float4 b, float area_b) {
    const float inter = box_inter_rn(a, b);
    return inter > 0.f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter)) : 0.f;
}

// Order-preserving map float -> uint32 (larger float => larger key; -0 < +0; NaN sorts by payload).
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// In-shared-memory bitonic sort of `n_pow2` 64-bit keys, DESCENDING. All threads of the block call it.
__device__ __forceinline__ void bitonic_sort_desc_u64(unsigned long long* keys, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool desc = ((i & k) == 0);
                    if (desc ? (a < b) : (a > b)) {
                        keys[i] = b;
                        keys[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace soswsod
