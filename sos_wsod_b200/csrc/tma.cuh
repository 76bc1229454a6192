// mbarrier + TMA (cp.async.bulk.tensor) PTX wrappers and the host-side tensor-map encoder, shared by the GEMM and
// the ROI-pool backward kernels.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace soswsod {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
// Low-latency spin for hand-offs between warps of one CTA (test_wait does not suspend the thread).
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2-D tensor map over a row-major [rows, cols] matrix with leading dimension ld (elements of elem_bytes);
// box = {box_cols (inner), box_rows}.  Out-of-bounds box elements are zero-filled.
// Encoded maps are cached per host thread (direct-mapped, 128 entries keyed on every argument): a training step re-uses
// the same few dozen (pointer, shape, box) combinations step after step -- the caching allocator hands the same
// blocks back -- so the driver call (~1-2 us each, two per GEMM / ROI-backward launch) is paid once.
struct TmapKey {
    const void* base;
    long long rows, cols, ld;
    int dtype, elem_bytes, box_cols, box_rows, swizzle;
    bool operator==(const TmapKey& o) const {
        return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && dtype == o.dtype &&
               elem_bytes == o.elem_bytes && box_cols == o.box_cols && box_rows == o.box_rows && swizzle == o.swizzle;
    }
};
struct TmapSlot {
    TmapKey key;
    CUtensorMap map;
    bool valid;
};
inline TmapSlot* tmap_cache() {
    static thread_local TmapSlot slots[128] = {};
    return slots;
}

inline int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* base, long long rows,
                        long long cols, long long ld, int box_cols, int box_rows, CUtensorMapSwizzle swizzle) {
    const TmapKey key{base, rows, cols, ld, (int)dtype, elem_bytes, box_cols, box_rows, (int)swizzle};
    unsigned long long hsh = reinterpret_cast<unsigned long long>(base) * 0x9E3779B97F4A7C15ull;
    hsh ^= (unsigned long long)rows * 0xBF58476D1CE4E5B9ull + (unsigned long long)cols * 0x94D049BB133111EBull +
           (unsigned long long)ld * 31ull + (unsigned long long)(box_cols * 131 + box_rows * 17 + (int)dtype * 7 + (int)swizzle);
    TmapSlot& slot = tmap_cache()[(hsh >> 32) & 127u];
    if (slot.valid && slot.key == key) {
        *map = slot.map;
        return SOSWSOD_OK;
    }
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return SOSWSOD_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dtype, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld box=%dx%d)", (int)r, rows, cols,
                  ld, box_cols, box_rows);
        return SOSWSOD_ERR_CUDA;
    }
    slot.key = key;
    slot.map = *map;
    slot.valid = true;
    return SOSWSOD_OK;
}

}  // namespace soswsod
