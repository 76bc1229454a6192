// ROI max-pool forward / backward for sm_100a (kernel (1) of the hot path).
//
// Forward: a CTA stages CT channel planes of one image in shared memory (fp32, exact), then its warps
// walk the ROIs of the CTA's chunk independently.  Inside a warp a lane owns one bin column of the ROI and
// loops over (channel, bin row) pairs (see roi_pool_fwd_kernel).
// Results of one ROI (CT*PH*PW values + argmax) are staged per warp and written out coalesced, as raw
// fp32 (bit-equal to torchvision), as argmax (int32 / uint16) and as the bf16 fc6 operand already
// multiplied by (objectness + 1).
//
// Backward: a CTA owns the gradient planes of a few channels of one image in shared memory, one consumer
// warp per plane (sole writer), fed by a TMA ring; a warp step takes the bins of one colour class of one
// roi, which are pairwise disjoint, so there is no atomic and no conflict detection anywhere and the
// summation order is fixed (deterministic).  See the comment above roi_pool_bwd_kernel.
//
// Bin arithmetic follows torchvision roi_pool exactly (restated in-repo by the reference at
// uwsod/projects/WSL/wsl/layers/csrc/ROILoopPool/ROILoopPool_cuda.cu:77-137): C round() of the fp32
// product, fp32 bin size, floor/ceil, clamp, strict '>' scan in row-major order, empty bin -> (0, -1).
#include "roi_plan.cuh"
#include "tma.cuh"

namespace soswsod {

constexpr int kFwdThreads = 512;
constexpr int kFwdWarps = kFwdThreads / 32;
constexpr int kMaxBins = 256;  // PH*PW limit

// One bin column of BWM cells over rows [hs, he): running row maximum (strict '>' keeps the FIRST best row).
// `row` points at (hs, ws).  A lane whose own range is one cell narrower (shortlane) masks the last load; that load
// is still inside the padded plane.
template <int BWM>
__device__ __forceinline__ void scan_rows(const float* __restrict__ row, int W, int hs, int he, bool shortlane,
                                          float& m, int& bh) {
#pragma unroll 2
    for (int h = hs; h < he; ++h, row += W) {
        float v[BWM];
#pragma unroll
        for (int j = 0; j < BWM; ++j) v[j] = row[j];
        if (shortlane) v[BWM - 1] = -FLT_MAX;
        float rm = v[0];
#pragma unroll
        for (int j = 1; j < BWM; ++j) rm = fmaxf(rm, v[j]);
        if (rm > m) {
            m = rm;
            bh = h;
        }
    }
}

constexpr int kFwdMaxP = 16;  // pooled_h, pooled_w limit of the staged kernel (larger grids use the global kernel)

// Lane mapping: lane = sub * PW + pw with nsub = 32 / PW sub-slots (PW = 7 -> 4 x 7 = 28 active lanes).  A lane owns
// one bin COLUMN pw of the ROI and walks the (channel, ph) pairs p = sub, sub + nsub, ... of the CTA's CT channels
// (c = p % CT, ph = p / CT), so its column range [ws, we) is loop-invariant and lanes of one step share ph.  Per
// bin it keeps only a running row maximum (one FMNMX per cell) and the first row that raised the maximum; the
// arg-max column is recovered by re-scanning that single row -- same result as the reference's strict '>' scan
// in row-major order (first maximum wins), NaNs never win, the stored value is the cell's own bit pattern.
// NT threads per CTA.  COMPACT (bf16 + uint16 outputs only -- the head engine's mode): results are staged as one
// 32-bit word per bin (bf16 value << 16 | uint16 argmax), which leaves room for 32 warps per SM.
template <int CT, int NT, bool COMPACT>
__global__ void __launch_bounds__(NT, 1)
roi_pool_fwd_kernel(const float* __restrict__ feat, int C, int H, int W, const float* __restrict__ rois, int R,
                    int PH, int PW, float scale, const float* __restrict__ row_scale, float row_scale_bias,
                    float* __restrict__ out_f32, int32_t* __restrict__ argmax_i32,
                    uint16_t* __restrict__ argmax_u16, __nv_bfloat16* __restrict__ out_bf16, long long ld_bf16,
                    int plane_stride, int rois_per_cta) {
    extern __shared__ __align__(16) float smem[];
    float* planes = smem;                                        // [CT][plane_stride]
    const int PP = PH * PW;
    constexpr int kWarps = NT / 32;
    float* stage_val = planes + CT * plane_stride;               // [kWarps][CT*PP] (COMPACT: packed words)
    int* stage_idx = reinterpret_cast<int*>(stage_val + kWarps * CT * PP);   // generic mode only
    int* bounds_all = COMPACT ? reinterpret_cast<int*>(stage_val + kWarps * CT * PP)
                              : stage_idx + kWarps * CT * PP;    // [kWarps][4*kFwdMaxP]

    const int groups = C / CT;
    const int b = blockIdx.x / groups;
    const int c0 = (blockIdx.x % groups) * CT;
    const int HW = H * W;

    // ---- stage CT planes (coalesced; 128-bit when the plane is 16B-aligned) ----
    {
        const float* src = feat + ((size_t)b * C + c0) * HW;
        if ((HW & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (plane_stride & 3) == 0) {
            const int n4 = HW >> 2;
            for (int c = 0; c < CT; ++c) {
                const float4* s4 = reinterpret_cast<const float4*>(src + (size_t)c * HW);
                float4* d4 = reinterpret_cast<float4*>(planes + c * plane_stride);
                for (int i = threadIdx.x; i < n4; i += NT) d4[i] = __ldg(s4 + i);
            }
        } else {
            for (int c = 0; c < CT; ++c)
                for (int i = threadIdx.x; i < HW; i += NT)
                    planes[c * plane_stride + i] = __ldg(src + (size_t)c * HW + i);
        }
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsub = 32 / PW;
    const int sub = lane / PW, pw = lane - sub * PW;
    const bool active = sub < nsub;
    float* sv = stage_val + warp * CT * PP;
    int* si = stage_idx + warp * CT * PP;
    int* bnd = bounds_all + warp * 4 * kFwdMaxP;
    const int n_out = CT * PP;
    const int npairs = CT * PH;

    const int r_begin = blockIdx.y * rois_per_cta;
    const int r_end = min(R, r_begin + rois_per_cta);
    for (int r = r_begin + warp; r < r_end; r += kWarps) {
        const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, PH, PW);
        if (g.batch != b) continue;  // warp-uniform
        if (lane < PH) {
            int hs = (int)floorf(__fmul_rn((float)lane, g.bin_h));
            int he = (int)ceilf(__fmul_rn((float)(lane + 1), g.bin_h));
            bnd[lane] = min(max(hs + g.rs_h, 0), H);
            bnd[kFwdMaxP + lane] = min(max(he + g.rs_h, 0), H);
        }
        if (lane >= 16 && lane - 16 < PW) {
            const int q = lane - 16;
            int ws = (int)floorf(__fmul_rn((float)q, g.bin_w));
            int we = (int)ceilf(__fmul_rn((float)(q + 1), g.bin_w));
            bnd[2 * kFwdMaxP + q] = min(max(ws + g.rs_w, 0), W);
            bnd[3 * kFwdMaxP + q] = min(max(we + g.rs_w, 0), W);
        }
        __syncwarp();
        // Column range of this lane; the widest range of the warp step picks ONE fully unrolled row scanner for
        // every lane (uniform branch).  Lanes one cell narrower mask their last load; anything narrower still
        // (bins clipped at the image border) takes the generic loop.
        int ws = 0, bw = 0;
        if (active) {
            ws = bnd[2 * kFwdMaxP + pw];
            bw = max(bnd[3 * kFwdMaxP + pw] - ws, 0);
        }
        const int bwmax = (int)__reduce_max_sync(FULL_MASK, (unsigned)bw);
        const bool fast = bw >= 1 && bw >= bwmax - 1 && bwmax <= 16;
        const float scl = (COMPACT && row_scale) ? __fadd_rn(row_scale[r], row_scale_bias) : 1.f;
        uint32_t* sp = reinterpret_cast<uint32_t*>(sv);
        const bool shortlane = bw < bwmax;
        if (active) {
            for (int p = sub; p < npairs; p += nsub) {
                const int c = p % CT, ph = p / CT;
                const int hs = bnd[ph], he = bnd[kFwdMaxP + ph];
                const float* pl = planes + c * plane_stride;
                float m = -FLT_MAX;
                int bh = -1;
                if (bw > 0 && he > hs) {
                    if (fast) {
                        const float* row = pl + hs * W + ws;
                        switch (bwmax) {
#define SOSWSOD_SCAN(BWM) case BWM: scan_rows<BWM>(row, W, hs, he, shortlane, m, bh); break;
                            SOSWSOD_SCAN(1) SOSWSOD_SCAN(2) SOSWSOD_SCAN(3) SOSWSOD_SCAN(4) SOSWSOD_SCAN(5) SOSWSOD_SCAN(6)
                            SOSWSOD_SCAN(7) SOSWSOD_SCAN(8) SOSWSOD_SCAN(9) SOSWSOD_SCAN(10) SOSWSOD_SCAN(11)
                            SOSWSOD_SCAN(12) SOSWSOD_SCAN(13) SOSWSOD_SCAN(14) SOSWSOD_SCAN(15) SOSWSOD_SCAN(16)
#undef SOSWSOD_SCAN
                        }
                    } else {
                        for (int h = hs; h < he; ++h) {
                            const float* row = pl + h * W + ws;
                            float rm = -FLT_MAX;
                            for (int j = 0; j < bw; ++j) rm = fmaxf(rm, row[j]);
                            if (rm > m) {
                                m = rm;
                                bh = h;
                            }
                        }
                    }
                }
                const bool empty = (he <= hs) || (bw <= 0);
                float outv = empty ? 0.f : -FLT_MAX;
                int idx = -1;
                if (bh >= 0) {
                    const float* row = pl + bh * W + ws;
                    int j = 0;
                    while (j < bw - 1 && row[j] != m) ++j;
                    outv = row[j];
                    idx = bh * W + ws + j;
                }
                if (COMPACT) {
                    const __nv_bfloat16 hv = __float2bfloat16_rn(__fmul_rn(outv, scl));
                    sp[c * PP + ph * PW + pw] = ((uint32_t)__bfloat16_as_ushort(hv) << 16) | (uint32_t)(idx < 0 ? 0xFFFF : idx);
                } else {
                    sv[c * PP + ph * PW + pw] = outv;
                    si[c * PP + ph * PW + pw] = idx;
                }
            }
        }
        __syncwarp();
        const size_t obase = ((size_t)r * C + c0) * PP;
        if (COMPACT) {
            // n_out = CT*PP words -> two outputs per lane and store where the destination is 4-byte aligned
            uint16_t* adst = argmax_u16 ? argmax_u16 + obase : nullptr;
            uint16_t* vdst = out_bf16 ? reinterpret_cast<uint16_t*>(out_bf16 + (size_t)r * ld_bf16 + (size_t)c0 * PP) : nullptr;
            const bool pair_ok = (n_out & 1) == 0 && (!adst || (reinterpret_cast<uintptr_t>(adst) & 3) == 0) &&
                                 (!vdst || (reinterpret_cast<uintptr_t>(vdst) & 3) == 0);
            if (pair_ok) {
                for (int i = lane; i < (n_out >> 1); i += 32) {
                    const uint32_t w0 = sp[2 * i], w1 = sp[2 * i + 1];
                    if (adst) reinterpret_cast<uint32_t*>(adst)[i] = (w0 & 0xFFFFu) | (w1 << 16);
                    if (vdst) reinterpret_cast<uint32_t*>(vdst)[i] = (w0 >> 16) | (w1 & 0xFFFF0000u);
                }
            } else {
                for (int i = lane; i < n_out; i += 32) {
                    const uint32_t w0 = sp[i];
                    if (adst) adst[i] = (uint16_t)(w0 & 0xFFFFu);
                    if (vdst) vdst[i] = (uint16_t)(w0 >> 16);
                }
            }
            __syncwarp();
            continue;
        }
        if (out_f32)
            for (int i = lane; i < n_out; i += 32) out_f32[obase + i] = sv[i];
        if (argmax_i32)
            for (int i = lane; i < n_out; i += 32) argmax_i32[obase + i] = si[i];
        if (argmax_u16)
            for (int i = lane; i < n_out; i += 32) argmax_u16[obase + i] = (uint16_t)(si[i] < 0 ? 0xFFFF : si[i]);
        if (out_bf16) {
            const float s = row_scale ? __fadd_rn(row_scale[r], row_scale_bias) : 1.f;
            __nv_bfloat16* dst = out_bf16 + (size_t)r * ld_bf16 + (size_t)c0 * PP;
            if ((n_out & 1) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
                __nv_bfloat162* d2 = reinterpret_cast<__nv_bfloat162*>(dst);
                for (int i = lane; i < (n_out >> 1); i += 32)
                    d2[i] = __floats2bfloat162_rn(__fmul_rn(sv[2 * i], s), __fmul_rn(sv[2 * i + 1], s));
            } else {
                for (int i = lane; i < n_out; i += 32) dst[i] = __float2bfloat16_rn(__fmul_rn(sv[i], s));
            }
        }
        __syncwarp();
    }
}

// Fallback for feature planes too large for shared memory: same arithmetic straight from global/L2,
// one warp per (roi, channel).
__global__ void roi_pool_fwd_global_kernel(const float* __restrict__ feat, int C, int H, int W,
                                           const float* __restrict__ rois, int R, int PH, int PW, float scale,
                                           const float* __restrict__ row_scale, float row_scale_bias,
                                           float* __restrict__ out_f32, int32_t* __restrict__ argmax_i32,
                                           uint16_t* __restrict__ argmax_u16,
                                           __nv_bfloat16* __restrict__ out_bf16, long long ld_bf16) {
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= (long long)R * C) return;
    const int r = (int)(gw / C), c = (int)(gw % C);
    const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, PH, PW);
    const float* pl = feat + ((size_t)g.batch * C + c) * H * W;
    const int PP = PH * PW;
    const float s = row_scale ? __fadd_rn(row_scale[r], row_scale_bias) : 1.f;
    for (int bin = lane; bin < PP; bin += 32) {
        const int ph = bin / PW, pw = bin % PW;
        int hs = (int)floorf(__fmul_rn((float)ph, g.bin_h));
        int he = (int)ceilf(__fmul_rn((float)(ph + 1), g.bin_h));
        hs = min(max(hs + g.rs_h, 0), H);
        he = min(max(he + g.rs_h, 0), H);
        int ws = (int)floorf(__fmul_rn((float)pw, g.bin_w));
        int we = (int)ceilf(__fmul_rn((float)(pw + 1), g.bin_w));
        ws = min(max(ws + g.rs_w, 0), W);
        we = min(max(we + g.rs_w, 0), W);
        const bool empty = (he <= hs) || (we <= ws);
        float m = empty ? 0.f : -FLT_MAX;
        int idx = -1;
        for (int h = hs; h < he; ++h)
            for (int w = ws; w < we; ++w) {
                const float v = __ldg(pl + h * W + w);
                if (v > m) {
                    m = v;
                    idx = h * W + w;
                }
            }
        const size_t o = ((size_t)r * C + c) * PP + bin;
        if (out_f32) out_f32[o] = m;
        if (argmax_i32) argmax_i32[o] = idx;
        if (argmax_u16) argmax_u16[o] = (uint16_t)(idx < 0 ? 0xFFFF : idx);
        if (out_bf16) out_bf16[(size_t)r * ld_bf16 + (size_t)c * PP + bin] = __float2bfloat16_rn(__fmul_rn(m, s));
    }
}

// ------------------------------------------------------------------------------------------------
// Backward
// ------------------------------------------------------------------------------------------------
// A CTA owns the gradient planes of CT consecutive channels of one image (optionally only a band of rows of
// them) in shared memory; consumer warp w is the ONLY writer of channel c0+w's plane, so there is no atomic and
// no cross-warp race, and the accumulation order is fixed (deterministic).  A producer warp streams the
// (roi, channel-group) tiles of argmax and grad_out through a TMA ring (cp.async.bulk.tensor + mbarriers): the
// 98-byte per-channel runs of the [R, C*PH*PW] matrices are neither 16-byte aligned nor long enough for wide
// loads; a TMA box only needs its first column 16-byte aligned, so each box starts at the aligned column at or
// below the group's first entry and the consumers add the remainder.
//
// Conflict freedom inside a warp step comes from PROPOSAL-BIN OWNERSHIP instead of run-time conflict detection:
// the arg-max of a bin lies inside the bin, and two bins of one roi can only share a cell when their row ranges
// AND column ranges overlap.  The producer computes, per roi, the smallest strides (mh, mw) such that bins mh rows
// (mw columns) apart are disjoint (2 x 2 for every roi at least 7 cells high and wide, larger for tiny rois whose
// bins repeat cells); a warp step then takes the bins of ONE colour class (ph % mh, pw % mw) of one roi -- pairwise
// disjoint by construction -- so every lane does a plain read-add-write on its own cell.  The argmax tensor must
// therefore be the one soswsod_roi_pool_forward produced for the same rois (the autograd contract).
constexpr int kBwdMaxCT = 8;
constexpr int kBwdMaxRT = 16;  // rois per ring stage (8 or 16)

struct BwdCfg {
    int CT, bands, band_rows, nbox, BW, stages, RT;
    int plane_stride;   // floats per plane band in smem
    size_t smem;
};

struct BwdRoiMeta {
    float scale;   // row_scale[r] + bias
    int code;      // 0 = skip this roi; else 1 | mh << 8 | mw << 16
};

template <typename GradT, typename ArgT, int PHT, int PWT>
__global__ void __launch_bounds__((kBwdMaxCT + 1) * 32, 1)
roi_pool_bwd_kernel(const __grid_constant__ CUtensorMap tmap_arg, const __grid_constant__ CUtensorMap tmap_grad,
                    const float* __restrict__ rois, int R, const float* __restrict__ row_scale, float row_scale_bias,
                    int N, int C, int H, int W, int PH_rt, int PW_rt, float spatial_scale,
                    float* __restrict__ grad_feat, BwdCfg cfg) {
    const int PH = PHT > 0 ? PHT : PH_rt;
    const int PW = PWT > 0 ? PWT : PW_rt;
    const int PP = PH * PW;
    extern __shared__ uint8_t smem_raw[];
    const int CT = cfg.CT, BW = cfg.BW, nbox = cfg.nbox, S = cfg.stages, RT = cfg.RT;
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    float* planes = reinterpret_cast<float*>(gen_base);                          // [CT][plane_stride]
    const uint32_t plane_bytes = (uint32_t)CT * cfg.plane_stride * 4;            // multiple of 128
    const uint32_t arg_box = (uint32_t)RT * BW * sizeof(ArgT);                   // multiple of 128
    const uint32_t grad_box = (uint32_t)RT * BW * sizeof(GradT);
    const uint32_t arg_stage = nbox * arg_box, grad_stage = nbox * grad_box;
    const uint32_t ring_off = plane_bytes;
    const uint32_t bar_off = ring_off + S * (arg_stage + grad_stage);
    auto full_bar = [&](int st) { return base + bar_off + 8u * st; };
    auto empty_bar = [&](int st) { return base + bar_off + 8u * (S + st); };
    int* s_range = reinterpret_cast<int*>(gen_base + bar_off + 16 * S);          // [2]
    BwdRoiMeta* s_meta = reinterpret_cast<BwdRoiMeta*>(s_range + 2);             // [S][kBwdMaxRT]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups = (C + CT - 1) / CT;
    int bid = blockIdx.x;
    const int band = bid % cfg.bands;
    bid /= cfg.bands;
    const int c0 = (bid % groups) * CT;
    const int b = bid / groups;
    const int HW = H * W;
    const int band_lo = min(band * cfg.band_rows, H) * W;
    const int band_hi = min((band + 1) * cfg.band_rows, H) * W;

    for (int i = threadIdx.x; i < CT * cfg.plane_stride; i += blockDim.x) planes[i] = 0.f;
    if (threadIdx.x == 0) {
        for (int st = 0; st < S; ++st) {
            mbar_init(full_bar(st), 1);
            mbar_init(empty_bar(st), CT);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_range[0] = N > 1 ? R : 0;
        s_range[1] = N > 1 ? 0 : R;
    }
    __syncthreads();
    if (N > 1) {  // rows of this image: [first, last+1) (exact when the rois are grouped by image, a superset otherwise)
        int lo = R, hi = 0;
        for (int r = threadIdx.x; r < R; r += blockDim.x)
            if ((int)rois[(size_t)r * 5] == b) {
                lo = min(lo, r);
                hi = max(hi, r + 1);
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(FULL_MASK, lo, o));
            hi = max(hi, __shfl_xor_sync(FULL_MASK, hi, o));
        }
        if (lane == 0) {
            atomicMin(&s_range[0], lo);   // integer bookkeeping only; gradients never go through an atomic
            atomicMax(&s_range[1], hi);
        }
        __syncthreads();
    }
    const int r_lo = s_range[0], r_hi = s_range[1];
    const int ntiles = r_hi > r_lo ? (r_hi - r_lo + RT - 1) / RT : 0;
    const int total_cols = C * PP;
    const int col0 = c0 * PP;
    const int col_a = col0 & ~(16 / (int)sizeof(ArgT) - 1);    // box start: 16-byte aligned column
    const int col_g = col0 & ~(16 / (int)sizeof(GradT) - 1);

    if (warp == CT) {
        // ---- producer warp: lane 0 drives the barriers and TMA, lanes < RT publish each roi's scale and colouring ----
        uint32_t tx = 0;
        for (int bx = 0; bx < nbox; ++bx) {
            if (col_a + bx * BW < total_cols) tx += arg_box;
            if (col_g + bx * BW < total_cols) tx += grad_box;
        }
        int st = 0;
        uint32_t phase = 0;
        for (int t = 0; t < ntiles; ++t) {
            const int r0 = r_lo + t * RT;
            // the roi geometry does not depend on the ring slot: compute it before waiting for the slot
            BwdRoiMeta meta;
            meta.scale = 0.f;
            meta.code = 0;
            if (lane < RT) {
                const int r = r0 + lane;
                if (r < r_hi && (N == 1 || (int)rois[(size_t)r * 5] == b)) {
                    const RoiGeom g = roi_geometry(rois + (size_t)r * 5, spatial_scale, PH, PW);
                    int mh = bin_disjoint_stride(g.bin_h, g.rs_h, PH, H);
                    int mw = bin_disjoint_stride(g.bin_w, g.rs_w, PW, W);
                    // a colour class must fit the 4 x 8 lane grid of a warp step
                    // (and strides of at least 2 keep the number of unrolled variants small)
                    mh = max(max(mh, 2), (PH + 3) / 4);
                    mw = max(max(mw, 2), (PW + 7) / 8);
                    meta.scale = row_scale ? __fadd_rn(row_scale[r], row_scale_bias) : 1.f;
                    meta.code = 1 | (mh << 8) | (mw << 16);
                }
            }
            mbar_wait(empty_bar(st), phase ^ 1u);
            if (lane < RT) s_meta[st * kBwdMaxRT + lane] = meta;
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(full_bar(st), tx);
                const uint32_t sa = base + ring_off + st * (arg_stage + grad_stage);
                const uint32_t sg = sa + arg_stage;
                for (int bx = 0; bx < nbox; ++bx) {
                    if (col_a + bx * BW < total_cols) tma_load_2d(sa + bx * arg_box, &tmap_arg, full_bar(st), col_a + bx * BW, r0);
                    if (col_g + bx * BW < total_cols) tma_load_2d(sg + bx * grad_box, &tmap_grad, full_bar(st), col_g + bx * BW, r0);
                }
            }
            if (++st == S) {
                st = 0;
                phase ^= 1u;
            }
        }
    } else if (warp < CT) {
        // this warp's plane as a shared-window address held in a register (opaque to the compiler, which otherwise
        // re-derives it from the kernel parameters inside the dependent read-add-write chain)
        uint32_t my_s = smem_u32(planes + warp * cfg.plane_stride);
        asm volatile("mov.u32 %0, %0;" : "+r"(my_s));
        const bool chan_ok = (c0 + warp) < C;
        const int ea0 = warp * PP + (col0 - col_a);
        const int eg0 = warp * PP + (col0 - col_g);
        const unsigned band_cells = (unsigned)(band_hi - band_lo);
        // Lane = (la, lb) = (lane >> 3, lane & 7) owns the block of mh x mw bins at (la*mh, lb*mw); colour step (i, j)
        // touches bin (la*mh + i, lb*mw + j) of every block.
        const int la = lane >> 3, lb = lane & 7;
        // byte offset, inside a ring stage, of this warp's entry `bin` of the roi in slot 0
        auto off_arg = [&](int bin) {
            int e = ea0 + bin, o = 0;
            if (e >= BW) { e -= BW; o = RT * BW; }
            return (o + e) * (int)sizeof(ArgT);
        };
        auto off_grad = [&](int bin) {
            int e = eg0 + bin, o = 0;
            if (e >= BW) { e -= BW; o = RT * BW; }
            return (o + e) * (int)sizeof(GradT);
        };
        // strides 2 x 2 (every roi at least PH x PW cells, the common case): the four operands of a lane sit at
        // offsets that never change, so they are computed once
        constexpr int kCode22 = 1 | (2 << 8) | (2 << 16);
        int oa2[4], og2[4];
        bool v2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ph = 2 * la + (k >> 1), pw = 2 * lb + (k & 1);
            v2[k] = ph < PH && pw < PW;
            oa2[k] = v2[k] ? off_arg(ph * PW + pw) : 0;
            og2[k] = v2[k] ? off_grad(ph * PW + pw) : 0;
        }
        int row_a = BW * (int)sizeof(ArgT), row_g = BW * (int)sizeof(GradT), rt = RT;
        asm volatile("mov.u32 %0, %0;" : "+r"(row_a));   // keep these in registers (see my_s)
        asm volatile("mov.u32 %0, %0;" : "+r"(row_g));
        asm volatile("mov.u32 %0, %0;" : "+r"(rt));
        int st = 0;
        uint32_t phase = 0;
        for (int t = 0; t < ntiles; ++t) {
            mbar_wait(full_bar(st), phase);
            if (chan_ok) {
                const uint8_t* pa = gen_base + ring_off + (size_t)st * (arg_stage + grad_stage);
                const uint8_t* pg = pa + arg_stage;
                const BwdRoiMeta* metas = s_meta + st * kBwdMaxRT;
                auto ld_arg = [&](int off) -> unsigned {
                    if (sizeof(ArgT) == 2) return (unsigned)*reinterpret_cast<const uint16_t*>(pa + off);
                    return *reinterpret_cast<const unsigned*>(pa + off);
                };
                auto ld_grad = [&](int off) -> float {
                    if (sizeof(GradT) == 2)
                        return __uint_as_float((unsigned)*reinterpret_cast<const uint16_t*>(pg + off) << 16);
                    return *reinterpret_cast<const float*>(pg + off);
                };
                // one colour step: plain read-add-write, lanes of a step never share a cell.  An empty bin (0xFFFF /
                // -1), an idle lane (0xFFFFFFFF) and a cell outside this CTA's row band all fail the range test.
                auto rmw = [&](unsigned a, float g) {
                    const unsigned rel = a - (unsigned)band_lo;
                    if (rel < band_cells) {
                        const uint32_t addr = my_s + rel * 4u;
                        float v;
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
                        v += g;
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
                    }
                    __syncwarp();   // colour classes of one roi may share cells: order the steps
                };
                // The tile is consumed as GROUPS of up to four colour steps: the operands of group g+1 are fetched
                // (8 independent loads) before the four ordered read-add-write steps of group g run, so the only
                // dependent chain left is the one on the plane.  A 2 x 2 roi is exactly one group; rois with larger
                // strides take several; a skipped roi (code 0) is an empty group.  (nrr, ni, nj) = next step to fetch.
                int nrr = 0, ni = 0, nj = 0;
                BwdRoiMeta m0 = metas[0];                                           // meta of roi nrr
                BwdRoiMeta m1 = metas[1];                                           // meta of roi nrr + 1 (RT >= 8)
                auto next_roi = [&]() {
                    ++nrr;
                    ni = nj = 0;
                    m0 = m1;
                    m1.code = 0;
                    if (nrr + 1 < rt) m1 = metas[nrr + 1];   // used one group later at the earliest
                };
                auto fetch_group = [&](unsigned (&a)[4], float (&g)[4]) {
                    const float sc = m0.scale;
                    if (m0.code == kCode22) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            a[k] = 0xFFFFFFFFu;
                            g[k] = 0.f;
                            if (v2[k]) {
                                a[k] = ld_arg(nrr * row_a + oa2[k]);
                                g[k] = ld_grad(nrr * row_g + og2[k]) * sc;
                            }
                        }
                        next_roi();
                        return;
                    }
                    const int mh = (m0.code >> 8) & 0xFF, mw = (m0.code >> 16) & 0xFF;   // 0 for a skipped roi
                    const int ph0 = la * mh, pw0 = lb * mw;
                    const int ra = nrr * row_a, rg = nrr * row_g;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        a[k] = 0xFFFFFFFFu;
                        g[k] = 0.f;
                        const int ph = ph0 + ni, pw = pw0 + nj;
                        if (ni < mh && ph < PH && pw < PW) {
                            a[k] = ld_arg(ra + off_arg(ph * PW + pw));
                            g[k] = ld_grad(rg + off_grad(ph * PW + pw)) * sc;
                        }
                        if (++nj >= mw) {
                            nj = 0;
                            ++ni;
                        }
                    }
                    if (ni >= mh) next_roi();
                };
                unsigned ca[4], na[4];
                float cg[4], ng[4];
                fetch_group(ca, cg);
                while (true) {
                    const bool more = nrr < rt;
                    if (more) fetch_group(na, ng);
#pragma unroll
                    for (int k = 0; k < 4; ++k) rmw(ca[k], cg[k]);
                    if (!more) break;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        ca[k] = na[k];
                        cg[k] = ng[k];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar(st));
            if (++st == S) {
                st = 0;
                phase ^= 1u;
            }
        }
    }
    __syncthreads();
    const int band_cells = band_hi - band_lo;
    for (int c = 0; c < CT && c0 + c < C; ++c) {
        float* dst = grad_feat + ((size_t)b * C + c0 + c) * HW + band_lo;
        const float* src = planes + c * cfg.plane_stride;
        for (int i = threadIdx.x; i < band_cells; i += blockDim.x) dst[i] = src[i];
    }
}

static int g_max_smem_optin = -1;
static int g_num_sms = -1;
static int query_device() {
    if (g_max_smem_optin >= 0) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    int v = 0, s = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&s, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    g_num_sms = s;
    g_max_smem_optin = v;
    return 0;
}
int device_num_sms() {
    query_device();
    return g_num_sms > 0 ? g_num_sms : 148;
}
int device_max_smem() {
    query_device();
    return g_max_smem_optin > 0 ? g_max_smem_optin : 227 * 1024;
}

template <int CT, int NT, bool COMPACT>
static int launch_fwd(const float* feat, int n, int c, int h, int w, const float* rois, int R, int PH, int PW,
                      float scale, const float* row_scale, float bias, float* out_f32, int32_t* a32, uint16_t* a16,
                      __nv_bfloat16* obf, long long ld, int plane_stride, size_t smem, cudaStream_t st) {
    const int groups = n * (c / CT);
    constexpr int kWarps = NT / 32;
    // Split the ROIs into `chunks` CTAs per channel group so that groups*chunks fills a whole number of
    // waves of the SMs (1 CTA/SM: the planes take most of shared memory) with the smallest tail.
    const int sms = device_num_sms();
    const int max_chunks = max(1, min(32, (R + kWarps - 1) / kWarps));
    int chunks = 1;
    double best = 1e30;
    for (int ch = 1; ch <= max_chunks; ++ch) {
        const int waves = (groups * ch + sms - 1) / sms;
        const double cost = (double)waves / ch;  // ~ time in units of "all ROIs of one group on one SM"
        if (cost < best - 1e-9) {
            best = cost;
            chunks = ch;
        }
    }
    const int rois_per_cta = (R + chunks - 1) / chunks;
    chunks = (R + rois_per_cta - 1) / rois_per_cta;
    auto kern = roi_pool_fwd_kernel<CT, NT, COMPACT>;
    SOSWSOD_ENSURE_SMEM(kern, smem);
    kern<<<dim3(groups, chunks), NT, smem, st>>>(feat, c, h, w, rois, R, PH, PW, scale, row_scale, bias, out_f32, a32, a16,
                                                 obf, ld, plane_stride, rois_per_cta);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

}  // namespace soswsod

using namespace soswsod;

extern "C" int soswsod_roi_pool_forward(const float* feat, int n, int c, int h, int w, const float* rois,
                                        int num_rois, int pooled_h, int pooled_w, float spatial_scale,
                                        const float* row_scale, float row_scale_bias, float* out_f32,
                                        void* argmax, int argmax_dtype, void* out_bf16, long long ld_bf16,
                                        const void* plan, size_t plan_bytes, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(n > 0 && c > 0 && h > 0 && w > 0 && num_rois >= 0, "roi_pool_forward: bad shape");
    if (num_rois == 0) return SOSWSOD_OK;  // empty proposal list: nothing to write
    SOSWSOD_CHECK_ARG(feat && rois, "roi_pool_forward: null feat/rois");
    SOSWSOD_CHECK_ARG(pooled_h > 0 && pooled_w > 0 && pooled_h * pooled_w <= kMaxBins,
                      "roi_pool_forward: pooled size %dx%d unsupported", pooled_h, pooled_w);
    SOSWSOD_CHECK_ARG(argmax_dtype == SOSWSOD_ARGMAX_I32 || argmax_dtype == SOSWSOD_ARGMAX_U16,
                      "roi_pool_forward: bad argmax dtype");
    SOSWSOD_CHECK_ARG(argmax_dtype == SOSWSOD_ARGMAX_I32 || (long long)h * w < 65535,
                      "roi_pool_forward: uint16 argmax needs h*w < 65535");
    SOSWSOD_CHECK_ARG(!out_bf16 || ld_bf16 >= (long long)c * pooled_h * pooled_w, "roi_pool_forward: ld_bf16 too small");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* a32 = argmax_dtype == SOSWSOD_ARGMAX_I32 ? (int32_t*)argmax : nullptr;
    uint16_t* a16 = argmax_dtype == SOSWSOD_ARGMAX_U16 ? (uint16_t*)argmax : nullptr;
    __nv_bfloat16* obf = (__nv_bfloat16*)out_bf16;
    const int PP = pooled_h * pooled_w;
    const int HW = h * w;
    const int max_smem = device_max_smem();
    if (plan) {
        SOSWSOD_CHECK_ARG(pooled_h == kPlanP && pooled_w == kPlanP, "roi_pool_forward: a plan is for 7x7 bins");
        if (plan_bytes < plan_total_bytes(num_rois)) {
            set_error("roi_pool_forward: plan of %zu bytes, need %zu", plan_bytes, plan_total_bytes(num_rois));
            return SOSWSOD_ERR_WORKSPACE;
        }
        if (!out_f32 && !a32) {   // operand mode: channel-interleaved planes + row-window table
            const int rc = launch_fwd_fast(feat, n, c, h, w, num_rois, plan, a16, obf, ld_bf16, st);
            if (rc != 0) return rc < 0 ? rc : SOSWSOD_OK;
        }
    }
    // pick the largest channel group whose planes + staging fit
    const int cts[4] = {8, 4, 2, 1};
    for (int k = 0; k < 4 && pooled_h <= kFwdMaxP && pooled_w <= kFwdMaxP; ++k) {
        const int CT = cts[k];
        if (c % CT) continue;
        const int nsub = 32 / pooled_w;
        // channel c sits c*(32/nsub) banks away from channel 0, so the nsub sub-slots of a warp step hit different banks
        int plane_stride = ((HW + 31) / 32) * 32 + ((32 / nsub) % 32);
        if (CT == 1) plane_stride = ((HW + 3) / 4) * 4 + 4;   // >= 1 float of padding after the last row
        const bool compact = !out_f32 && !a32;   // bf16 (+ uint16 argmax) outputs only: 32 warps, packed staging
        const int warps = compact ? 32 : kFwdWarps;
        const size_t smem = (size_t)CT * plane_stride * 4 + (size_t)warps * CT * PP * (compact ? 4 : 8) +
                            (size_t)warps * 4 * kFwdMaxP * 4;
        if (smem > (size_t)max_smem) continue;
#define SOSWSOD_FWD(CTV)                                                                                              \
    case CTV:                                                                                                         \
        return compact ? launch_fwd<CTV, 1024, true>(feat, n, c, h, w, rois, num_rois, pooled_h, pooled_w, spatial_scale, \
                                                     row_scale, row_scale_bias, out_f32, a32, a16, obf, ld_bf16,          \
                                                     plane_stride, smem, st)                                              \
                       : launch_fwd<CTV, kFwdThreads, false>(feat, n, c, h, w, rois, num_rois, pooled_h, pooled_w,        \
                                                             spatial_scale, row_scale, row_scale_bias, out_f32, a32, a16, \
                                                             obf, ld_bf16, plane_stride, smem, st);
        switch (CT) {
            SOSWSOD_FWD(8)
            SOSWSOD_FWD(4)
            SOSWSOD_FWD(2)
            SOSWSOD_FWD(1)
        }
#undef SOSWSOD_FWD
    }
    // plane larger than shared memory: global-memory kernel
    const long long warps = (long long)num_rois * c;
    const int threads = 256;
    const long long blocks = (warps * 32 + threads - 1) / threads;
    roi_pool_fwd_global_kernel<<<(unsigned)blocks, threads, 0, st>>>(feat, c, h, w, rois, num_rois, pooled_h,
                                                                    pooled_w, spatial_scale, row_scale,
                                                                    row_scale_bias, out_f32, a32, a16, obf, ld_bf16);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

static bool pick_bwd_cfg(int n, int c, int h, int w, int PP, int arg_bytes, int grad_bytes, BwdCfg* out) {
    const int max_smem = device_max_smem();
    const int sms = device_num_sms();
    const int align_elems = 16 / (arg_bytes < grad_bytes ? arg_bytes : grad_bytes);
    bool found = false;
    double best_cost = 0;
    for (int bands = 1; bands <= 64; ++bands) {
        const int band_rows = (h + bands - 1) / bands;
        if (bands > 1 && (long long)(bands - 1) * band_rows >= h) continue;  // empty last band
        const int plane_stride = ((band_rows * w + 31) / 32) * 32;
        for (int CT = kBwdMaxCT; CT >= 1; --CT) {
            if (CT > c) continue;
            const int cols = CT * PP + align_elems - 1;   // + the alignment remainder of the first column
            const int nbox = (cols + 247) / 248;
            const int BW = (((cols + nbox - 1) / nbox) + 7) / 8 * 8;
            if (BW > 256 || nbox > 2) continue;
            const size_t fixed = (size_t)CT * plane_stride * 4 + 128 /*align*/ + 1024 /*barriers, roi meta*/;
            for (int RT = kBwdMaxRT; RT >= 8; RT >>= 1) {
                const size_t stage = (size_t)nbox * RT * BW * (arg_bytes + grad_bytes);
                if (fixed + 2 * stage > (size_t)max_smem) continue;
                int stages = (int)(((size_t)max_smem - fixed) / stage);
                if (stages > 4) stages = 4;
                const long long ctas = (long long)n * ((c + CT - 1) / CT) * bands;
                const long long waves = (ctas + sms - 1) / sms;
                // every CTA streams all rois of its image once: time ~ waves, with a mild preference for deeper rings
                const double cost = (double)waves + (RT == 8 ? 0.05 : 0.0) + (stages < 3 ? 0.02 : 0.0);
                if (!found || cost < best_cost - 1e-9) {
                    found = true;
                    best_cost = cost;
                    out->CT = CT;
                    out->bands = bands;
                    out->band_rows = band_rows;
                    out->nbox = nbox;
                    out->BW = BW;
                    out->stages = stages;
                    out->RT = RT;
                    out->plane_stride = plane_stride;
                    out->smem = fixed + (size_t)stages * stage;
                }
                break;  // the largest RT that fits is enough for this (bands, CT)
            }
        }
        if (found && best_cost < 1.5) break;
    }
    return found;
}

template <typename GradT, typename ArgT>
static int launch_bwd(const void* grad, long long ld_grad, const void* argmax, const float* rois, int R,
                      const float* row_scale, float bias, int n, int c, int h, int w, int PH, int PW, float scale,
                      float* grad_feat, cudaStream_t st) {
    const int HW = h * w;
    const int PP = PH * PW;
    BwdCfg cfg;
    const bool aligned = ((uintptr_t)grad & 15) == 0 && ((uintptr_t)argmax & 15) == 0 &&
                         ((ld_grad * (long long)sizeof(GradT)) & 15) == 0 && (((long long)c * PP * sizeof(ArgT)) & 15) == 0;
    if (aligned && pick_bwd_cfg(n, c, h, w, PP, (int)sizeof(ArgT), (int)sizeof(GradT), &cfg)) {
        CUtensorMap ta, tg;
        int rc = make_tmap_2d(&ta, sizeof(ArgT) == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_INT32,
                              (int)sizeof(ArgT), argmax, R, (long long)c * PP, (long long)c * PP, cfg.BW, cfg.RT,
                              CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
        rc = make_tmap_2d(&tg, sizeof(GradT) == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                          (int)sizeof(GradT), grad, R, (long long)c * PP, ld_grad, cfg.BW, cfg.RT, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
        const int grid = n * ((c + cfg.CT - 1) / cfg.CT) * cfg.bands;
        const int threads = (cfg.CT + 1) * 32;
        if (PH == 7 && PW == 7) {
            auto kern = roi_pool_bwd_kernel<GradT, ArgT, 7, 7>;
            SOSWSOD_ENSURE_SMEM(kern, cfg.smem);
            kern<<<grid, threads, cfg.smem, st>>>(ta, tg, rois, R, row_scale, bias, n, c, h, w, PH, PW, scale, grad_feat, cfg);
        } else {
            auto kern = roi_pool_bwd_kernel<GradT, ArgT, 0, 0>;
            SOSWSOD_ENSURE_SMEM(kern, cfg.smem);
            kern<<<grid, threads, cfg.smem, st>>>(ta, tg, rois, R, row_scale, bias, n, c, h, w, PH, PW, scale, grad_feat, cfg);
        }
        SOSWSOD_CHECK_LAUNCH();
        return SOSWSOD_OK;
    }
    // There is no atomic path.  The TMA-fed kernels need 16-byte aligned rows (the Python shim pads the channel count
    // of odd layouts, ops.roi_pool_backward) and a band of a plane must fit shared memory.
    set_error("roi_pool_backward: unsupported layout (grad / arg-max rows must be 16-byte aligned: pad the channel count to a "
              "multiple of 8; planes of %d x %d cells must fit shared memory in at most 64 row bands)", h, w);
    return SOSWSOD_ERR_UNSUPPORTED;
}

extern "C" int soswsod_roi_pool_backward(const void* grad_out, int grad_dtype, long long ld_grad,
                                         const void* argmax, int argmax_dtype, const float* rois, int num_rois,
                                         const float* row_scale, float row_scale_bias, int n, int c, int h, int w,
                                         int pooled_h, int pooled_w, float spatial_scale, float* grad_feat,
                                         const void* plan, size_t plan_bytes, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(grad_out && argmax && rois && grad_feat, "roi_pool_backward: null pointer");
    SOSWSOD_CHECK_ARG(n > 0 && c > 0 && h > 0 && w > 0 && num_rois >= 0, "roi_pool_backward: bad shape");
    SOSWSOD_CHECK_ARG(grad_dtype == SOSWSOD_DTYPE_F32 || grad_dtype == SOSWSOD_DTYPE_BF16, "roi_pool_backward: bad grad dtype");
    SOSWSOD_CHECK_ARG(argmax_dtype == SOSWSOD_ARGMAX_I32 || argmax_dtype == SOSWSOD_ARGMAX_U16, "roi_pool_backward: bad argmax dtype");
    const int PP = pooled_h * pooled_w;
    SOSWSOD_CHECK_ARG(ld_grad >= (long long)c * PP, "roi_pool_backward: ld_grad too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (plan && num_rois > 0) {
        SOSWSOD_CHECK_ARG(pooled_h == kPlanP && pooled_w == kPlanP, "roi_pool_backward: a plan is for 7x7 bins");
        if (plan_bytes < plan_total_bytes(num_rois)) {
            set_error("roi_pool_backward: plan of %zu bytes, need %zu", plan_bytes, plan_total_bytes(num_rois));
            return SOSWSOD_ERR_WORKSPACE;
        }
        if (argmax_dtype == SOSWSOD_ARGMAX_U16) {
            const int rc = launch_bwd_fast(grad_out, grad_dtype, ld_grad, (const uint16_t*)argmax, num_rois, plan, n, c, h, w,
                                           grad_feat, st);
            if (rc != 0) return rc < 0 ? rc : SOSWSOD_OK;
        }
    }
    if (grad_dtype == SOSWSOD_DTYPE_F32) {
        if (argmax_dtype == SOSWSOD_ARGMAX_I32)
            return launch_bwd<float, int32_t>(grad_out, ld_grad, argmax, rois, num_rois, row_scale, row_scale_bias, n, c, h, w, pooled_h, pooled_w, spatial_scale, grad_feat, st);
        return launch_bwd<float, uint16_t>(grad_out, ld_grad, argmax, rois, num_rois, row_scale, row_scale_bias, n, c, h, w, pooled_h, pooled_w, spatial_scale, grad_feat, st);
    }
    if (argmax_dtype == SOSWSOD_ARGMAX_I32)
        return launch_bwd<__nv_bfloat16, int32_t>(grad_out, ld_grad, argmax, rois, num_rois, row_scale, row_scale_bias, n, c, h, w, pooled_h, pooled_w, spatial_scale, grad_feat, st);
    return launch_bwd<__nv_bfloat16, uint16_t>(grad_out, ld_grad, argmax, rois, num_rois, row_scale, row_scale_bias, n, c, h, w, pooled_h, pooled_w, spatial_scale, grad_feat, st);
}

extern "C" size_t soswsod_roi_pool_plan_bytes(int num_rois, int pooled_h, int pooled_w) {
    if (pooled_h != kPlanP || pooled_w != kPlanP || num_rois <= 0) return 0;
    return plan_total_bytes(num_rois);
}

extern "C" int soswsod_roi_pool_plan(const float* rois, int num_rois, int n, int h, int w, int pooled_h, int pooled_w,
                                     float spatial_scale, const float* row_scale, float row_scale_bias, void* plan,
                                     size_t plan_bytes, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(pooled_h == kPlanP && pooled_w == kPlanP, "roi_pool_plan: only 7x7 bins have a plan");
    SOSWSOD_CHECK_ARG(rois && plan && num_rois > 0, "roi_pool_plan: null pointer / no rois");
    SOSWSOD_CHECK_ARG(n > 0 && n <= kPlanMaxImages, "roi_pool_plan: 1..%d images per call", kPlanMaxImages);
    SOSWSOD_CHECK_ARG(h > 0 && w > 0 && h < 65536 && w < 65536, "roi_pool_plan: bad plane size");
    SOSWSOD_CHECK_ARG((reinterpret_cast<uintptr_t>(plan) & 127) == 0, "roi_pool_plan: plan must be 128-byte aligned");
    if (plan_bytes < plan_total_bytes(num_rois)) {
        set_error("roi_pool_plan: plan of %zu bytes, need %zu", plan_bytes, plan_total_bytes(num_rois));
        return SOSWSOD_ERR_WORKSPACE;
    }
    return launch_roi_plan(rois, num_rois, n, h, w, spatial_scale, row_scale, row_scale_bias, plan, (cudaStream_t)stream);
}
