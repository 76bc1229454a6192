// ROI max-pool, 7 x 7 bins, the head engine's operand mode (bf16 value * (objectness + 1) and uint16 arg-max):
// the fast forward / backward kernels behind soswsod_roi_pool_forward / _backward when a plan is supplied.
//
//   roi_plan_kernel        per call: groups the rois by image and stores, per roi, its 7 + 7 bin bounds, its scale
//                          factor and the backward's colour strides (64 bytes) -- bin arithmetic is done ONCE per roi
//                          instead of once per (roi, channel group) and per direction.
//   roi_pool_fwd_fast      channel-interleaved planes in shared memory ([cell][CI] fp32, one 128-bit load = one cell
//                          of 4 channels) plus a table of 1 x 4 row-window first-maxima of the same layout and a byte
//                          table of their offsets: a bin row of up to 8 cells costs two loads, the arg-max is the
//                          winning window's cell plus one byte lookup.  Exactly torchvision's result: strict '>' in
//                          row-major order (first maximum wins), empty bin -> (0, -1), the stored value is the
//                          cell's own bit pattern.
//   roi_pool_bwd_fast      see the comment above the kernel.
#include <stdlib.h>

#include "roi_plan.cuh"
#include "tma.cuh"

namespace soswsod {

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
constexpr int kPlanThreads = 1024;

__global__ void __launch_bounds__(kPlanThreads)
roi_plan_kernel(const float* __restrict__ rois, int R, int n, int H, int W, float spatial_scale,
                const float* __restrict__ row_scale, float row_scale_bias, uint8_t* __restrict__ plan) {
    __shared__ int s_warp[32];
    __shared__ int s_base[2];
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* img_start = reinterpret_cast<int*>(plan);
    int* order = reinterpret_cast<int*>(plan + kPlanHeaderBytes);
    RoiRecord* rec = reinterpret_cast<RoiRecord*>(plan + plan_records_offset(R));

    // one CTA per (image, chunk of 1024 rois): position of the chunk's first roi of this image inside `order` =
    // rois of earlier images + rois of this image in earlier chunks
    const int r_chunk = blockIdx.y * kPlanThreads;
    int before = 0, mine = 0;
    for (int r = threadIdx.x; r < R; r += kPlanThreads) {
        const int rb = (int)rois[(size_t)r * 5];
        before += ((rb >= 0 && rb < b) || (rb == b && r < r_chunk)) ? 1 : 0;
        mine += (rb == b) ? 1 : 0;
    }
    before = warp_sum_int(before);
    mine = warp_sum_int(mine);
    if (threadIdx.x < 2) s_base[threadIdx.x] = 0;
    __syncthreads();
    if (lane == 0) {
        atomicAdd(&s_base[0], before);   // integer bookkeeping
        atomicAdd(&s_base[1], mine);
    }
    __syncthreads();
    const int base = s_base[0];
    if (threadIdx.x == 0 && blockIdx.y == 0) {
        img_start[b] = base;
        if (b == n - 1) img_start[n] = base + s_base[1];
    }
    // stable compaction of this chunk's rois of image b + their records
    {
        const int r = r_chunk + threadIdx.x;
        bool flag = false;
        if (r < R) flag = ((int)rois[(size_t)r * 5] == b);
        const unsigned bal = __ballot_sync(FULL_MASK, flag);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int wpre = 0;
        for (int i = 0; i < warp; ++i) wpre += s_warp[i];
        if (flag) {
            order[base + wpre + __popc(bal & ((1u << lane) - 1u))] = r;
            const RoiGeom g = roi_geometry(rois + (size_t)r * 5, spatial_scale, kPlanP, kPlanP);
            RoiRecord rc;
#pragma unroll
            for (int p = 0; p < kPlanP; ++p) {
                rc.hb[p][0] = (uint16_t)bin_start(p, g.bin_h, g.rs_h, H);
                rc.hb[p][1] = (uint16_t)bin_end(p, g.bin_h, g.rs_h, H);
                rc.wb[p][0] = (uint16_t)bin_start(p, g.bin_w, g.rs_w, W);
                rc.wb[p][1] = (uint16_t)bin_end(p, g.bin_w, g.rs_w, W);
            }
            rc.mh = (uint8_t)max(bin_disjoint_stride(g.bin_h, g.rs_h, kPlanP, H), 2);
            rc.mw = (uint8_t)max(bin_disjoint_stride(g.bin_w, g.rs_w, kPlanP, W), 2);
            rc.batch = (uint16_t)b;
            rc.scale = row_scale ? __fadd_rn(row_scale[r], row_scale_bias) : 1.f;
            const uint4* s4 = reinterpret_cast<const uint4*>(&rc);
            uint4* d4 = reinterpret_cast<uint4*>(rec + r);
#pragma unroll
            for (int i = 0; i < 4; ++i) d4[i] = s4[i];
        }
        if (b == 0 && r < R) {   // rois of no image: in no list, and a record no CTA of the backward matches
            const int rb = (int)rois[(size_t)r * 5];
            if (rb < 0 || rb >= n) {
                uint4* d4 = reinterpret_cast<uint4*>(rec + r);
                d4[0] = d4[1] = d4[2] = make_uint4(0u, 0u, 0u, 0u);
                d4[3] = make_uint4(0u, 0u, 0xFFFF0202u, 0u);
            }
        }
    }
}

int launch_roi_plan(const float* rois, int R, int n, int h, int w, float scale, const float* row_scale, float bias,
                    void* plan, cudaStream_t st) {
    roi_plan_kernel<<<dim3(n, (R + kPlanThreads - 1) / kPlanThreads), kPlanThreads, 0, st>>>(rois, R, n, h, w, scale, row_scale, bias, static_cast<uint8_t*>(plan));
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
constexpr int kFastThreads = 1024;
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kPP = kPlanP * kPlanP;

template <int CI>
__device__ __forceinline__ void ld_cell(float (&v)[CI], const float* p) {
    if constexpr (CI == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (CI == 2) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = *p;
    }
}
template <int CI>
__device__ __forceinline__ void st_cell(float* p, const float (&v)[CI]) {
    if constexpr (CI == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (CI == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
        *p = v[0];
    }
}

__device__ __forceinline__ uint32_t pack_out(float val, float scl, int idx) {
    const __nv_bfloat16 hv = __float2bfloat16_rn(__fmul_rn(val, scl));
    return ((uint32_t)__bfloat16_as_ushort(hv) << 16) | (uint32_t)(idx < 0 ? 0xFFFF : idx);
}

// One bin per lane, CI channels at once.  NL loads per bin row: DIRECT reads NL cells of the plane (bins at most 4
// wide), otherwise NL windows of the 1 x 4 table (bins 4 .. 4*NL wide); `off` are the cell offsets of the loads
// inside the bin row (clamped into the lane's own bin, so a narrower lane re-reads a cell instead of masking; a
// re-read never wins the strict '>').  Per channel the lane keeps the running maximum and the cell index of the
// load that raised it; for a window that is the window's first cell, and the byte table `otab` ([cell][CI]) gives
// the offset of the first maximum inside the window.  Table entries carry the bit pattern of that first maximum
// (or -FLT_MAX when no cell of the window beats -FLT_MAX), so value and index equal torchvision's scan exactly.
template <int CI, int NL, bool DIRECT>
__device__ __forceinline__ void pool_bin(const float* __restrict__ xs, const float* __restrict__ tab,
                                         const uint8_t* __restrict__ otab, int W, int hs, int nrows, int nrmax, int ws,
                                         const int (&off)[NL], float scl, uint32_t* __restrict__ sp) {
    float m[CI];
    int pos[CI];
#pragma unroll
    for (int c = 0; c < CI; ++c) {
        m[c] = -FLT_MAX;
        pos[c] = -1;
    }
    int rowcell = hs * W + ws;
    const float* src = (DIRECT ? xs : tab) + rowcell * CI;
    const int rstride = W * CI;
    for (int h = 0; h < nrmax; ++h, src += rstride, rowcell += W) {
        if (h < nrows) {
#pragma unroll
            for (int t = 0; t < NL; ++t) {
                float v[CI];
                ld_cell<CI>(v, src + off[t] * CI);
                const int wc = rowcell + off[t];
#pragma unroll
                for (int c = 0; c < CI; ++c)
                    if (v[c] > m[c]) {
                        m[c] = v[c];
                        pos[c] = wc;
                    }
            }
        }
    }
    const bool empty = nrows <= 0;   // bw == 0 lanes arrive with nrows = 0
#pragma unroll
    for (int c = 0; c < CI; ++c) {
        float outv = empty ? 0.f : -FLT_MAX;
        int idx = -1;
        if (pos[c] >= 0) {
            outv = m[c];
            idx = pos[c];
            if (!DIRECT) idx += otab[pos[c] * CI + c];
        }
        sp[c * kPP] = pack_out(outv, scl, idx);
    }
}

// Any bin shape (bins clipped at the plane border, bins wider than 16 cells): plain scan, one bin per lane.
template <int CI>
__device__ __forceinline__ void pool_bin_generic(const float* __restrict__ xs, int W, int hs, int he, int ws, int we,
                                                 float scl, uint32_t* __restrict__ sp) {
    const bool empty = he <= hs || we <= ws;
#pragma unroll
    for (int c = 0; c < CI; ++c) {
        float m = -FLT_MAX;
        int idx = -1;
        for (int h = hs; h < he; ++h)
            for (int w = ws; w < we; ++w) {
                const float v = xs[(h * W + w) * CI + c];
                if (v > m) {
                    m = v;
                    idx = h * W + w;
                }
            }
        sp[c * kPP] = pack_out(empty ? 0.f : m, scl, idx);
    }
}

template <int CI>
__global__ void __launch_bounds__(kFastThreads, 1)
roi_pool_fwd_fast_kernel(const float* __restrict__ feat, int C, int H, int W, const int* __restrict__ img_start,
                         const int* __restrict__ order, const RoiRecord* __restrict__ rec,
                         uint16_t* __restrict__ argmax_u16, __nv_bfloat16* __restrict__ out_bf16, long long ld_bf16,
                         int cells_pad, int chunks) {
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                                   // [cells_pad][CI]
    float* tab = xs + (size_t)cells_pad * CI;           // [cells_pad][CI]  first maximum of cells i .. i+3
    uint32_t* stage = reinterpret_cast<uint32_t*>(tab + (size_t)cells_pad * CI);   // [warps][CI * 49]
    uint8_t* otab = reinterpret_cast<uint8_t*>(stage + kFastWarps * CI * kPP);     // [cells_pad][CI]  its offset 0..3
    const int HW = H * W;
    const int groups = C / CI;
    const int b = blockIdx.x / groups;
    const int c0 = (blockIdx.x % groups) * CI;

    // ---- stage the CI planes channel-interleaved, then the window table ----
    {
        const float* src = feat + ((size_t)b * C + c0) * HW;
        for (int i = threadIdx.x; i < cells_pad; i += kFastThreads) {
            float v[CI];
#pragma unroll
            for (int c = 0; c < CI; ++c) v[c] = i < HW ? __ldg(src + (size_t)c * HW + i) : 0.f;
            st_cell<CI>(xs + i * CI, v);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < cells_pad; i += kFastThreads) {
            float v[CI];
            uint32_t o = 0;
#pragma unroll
            for (int c = 0; c < CI; ++c) v[c] = -FLT_MAX;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float u[CI];
                ld_cell<CI>(u, xs + min(i + k, cells_pad - 1) * CI);
#pragma unroll
                for (int c = 0; c < CI; ++c)
                    if (u[c] > v[c]) {
                        v[c] = u[c];
                        o = (o & ~(0xFFu << (8 * c))) | ((uint32_t)k << (8 * c));
                    }
            }
            st_cell<CI>(tab + i * CI, v);
            if constexpr (CI == 4) reinterpret_cast<uint32_t*>(otab)[i] = o;
            else if constexpr (CI == 2) reinterpret_cast<uint16_t*>(otab)[i] = (uint16_t)o;
            else otab[i] = (uint8_t)o;
        }
        __syncthreads();
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / kPlanP, pw = lane - sub * kPlanP;
    const bool active = lane < 4 * kPlanP;
    uint32_t* sp = stage + warp * (CI * kPP);
    constexpr int n_out = CI * kPP;

    const int seg_lo = img_start[b], seg_hi = img_start[b + 1];
    const int per = (seg_hi - seg_lo + chunks - 1) / chunks;
    const int lo = seg_lo + blockIdx.y * per;
    const int hi = min(lo + per, seg_hi);
    for (int i = lo + warp; i < hi; i += kFastWarps) {
        const int r = __ldg(order + i);
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(rec + r);
        uint32_t hb0 = 0, hb1 = 0, wb = 0;
        if (active) {
            hb0 = __ldg(rw + sub);
            if (sub + 4 < kPlanP) hb1 = __ldg(rw + sub + 4);
            wb = __ldg(rw + kPlanP + pw);
        }
        const float scl = __uint_as_float(__ldg(rw + 15));
        const int ws = (int)(wb & 0xFFFFu), we = (int)(wb >> 16);
        const int bw = active ? we - ws : 0;
        const int bwmax = (int)__reduce_max_sync(FULL_MASK, (unsigned)bw);
        const int bwmin = (int)__reduce_min_sync(FULL_MASK, active ? (unsigned)bw : 0xFFFFu);
        // 0: direct (bins <= 4 wide), 1: window table, 2: generic
        const int mode = bwmax <= 4 ? 0 : ((bwmax <= 16 && bwmin >= 4) ? 1 : 2);
        const int nl = mode == 0 ? max(bwmax, 1) : (bwmax + 3) >> 2;
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            const int ph = sub + 4 * s;
            const bool on = active && ph < kPlanP;
            const uint32_t hb = s ? hb1 : hb0;
            const int hs = (int)(hb & 0xFFFFu), he = (int)(hb >> 16);
            const int nrows = (on && bw > 0) ? he - hs : 0;
            const int nrmax = (int)__reduce_max_sync(FULL_MASK, (unsigned)nrows);
            uint32_t* spl = sp + ph * kPlanP + pw;
            if (mode == 2) {
                if (on) pool_bin_generic<CI>(xs, W, hs, he, ws, we, scl, spl);
                continue;
            }
            if (!on) {
                // idle lanes still walk the (uniform) row loop with nrows = 0; they must not write
                continue;
            }
#define SOSWSOD_DIRECT(NLV)                                                                       \
    case NLV: {                                                                                   \
        int off[NLV];                                                                             \
        _Pragma("unroll") for (int t = 0; t < NLV; ++t) off[t] = max(min(t, bw - 1), 0);           \
        pool_bin<CI, NLV, true>(xs, tab, otab, W, hs, nrows, nrmax, ws, off, scl, spl);             \
    } break;
#define SOSWSOD_TABLE(NLV)                                                                        \
    case NLV: {                                                                                   \
        int off[NLV];                                                                             \
        _Pragma("unroll") for (int t = 0; t < NLV; ++t) off[t] = min(4 * t, bw - 4);              \
        pool_bin<CI, NLV, false>(xs, tab, otab, W, hs, nrows, nrmax, ws, off, scl, spl);            \
    } break;
            if (mode == 0) {
                switch (nl) {
                    SOSWSOD_DIRECT(1) SOSWSOD_DIRECT(2) SOSWSOD_DIRECT(3) SOSWSOD_DIRECT(4)
                }
            } else {
                switch (nl) {
                    SOSWSOD_TABLE(2) SOSWSOD_TABLE(3) SOSWSOD_TABLE(4)
                }
            }
#undef SOSWSOD_DIRECT
#undef SOSWSOD_TABLE
        }
        __syncwarp();
        // ---- coalesced write-out of the roi's CI*49 (value, arg-max) pairs ----
        uint16_t* adst = argmax_u16 + ((size_t)r * C + c0) * kPP;
        uint16_t* vdst = reinterpret_cast<uint16_t*>(out_bf16 + (size_t)r * ld_bf16 + (size_t)c0 * kPP);
        if (CI == 4) {
            for (int q = lane; q < n_out / 4; q += 32) {
                const uint4 w4 = *reinterpret_cast<const uint4*>(sp + 4 * q);
                uint2 a, v;
                a.x = (w4.x & 0xFFFFu) | (w4.y << 16);
                a.y = (w4.z & 0xFFFFu) | (w4.w << 16);
                v.x = (w4.x >> 16) | (w4.y & 0xFFFF0000u);
                v.y = (w4.z >> 16) | (w4.w & 0xFFFF0000u);
                reinterpret_cast<uint2*>(adst)[q] = a;
                reinterpret_cast<uint2*>(vdst)[q] = v;
            }
        } else {
            for (int q = lane; q < n_out / 2; q += 32) {
                const uint2 w2 = *reinterpret_cast<const uint2*>(sp + 2 * q);
                reinterpret_cast<uint32_t*>(adst)[q] = (w2.x & 0xFFFFu) | (w2.y << 16);
                reinterpret_cast<uint32_t*>(vdst)[q] = (w2.x >> 16) | (w2.y & 0xFFFF0000u);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// forward, "half table" layout: maps whose two interleaved planes + full window table do not fit one SM's shared
// memory (96x128, 108x144: the large test-time views).  The 1 x 4 window table is kept for EVEN window starts only
// and lives next to its plane row: a row is [W cells][W/2 window entries] (pitch 1.5 W cells, W even), so one row
// pointer serves both and a load is named by its COLUMN in that row (col < W: the cell itself, col >= W: the window
// starting at cell 2 (col - W)).  A bin row [ws, we) is covered, left to right, by
//     [cell ws if ws is odd]  windows at even starts a, a+4, .., b (b = the last even start with b+4 <= we)  [cell we-1]
// (no window fits a 4-wide bin at an odd start: its cells are read one by one); lanes with a shorter list repeat their
// last load.  Loads only ever cover cells of the bin and are visited left to right, and a re-read never wins the strict
// '>': value and arg-max are torchvision's, as in the full-table kernel.  The winner is recorded as (row << 8 | col).
// (Four channels at a time never pay here: wherever they fit with the half table, two fit with the full table, which
// needs a third fewer loads per bin row -- measured 687 vs 708 us at 72x96.)
struct BinCols {
    int mode, ws, we, W, hd, a, bb, nw, cend;
    __device__ __forceinline__ BinCols(int mode_, int ws_, int we_, int W_) : mode(mode_), ws(ws_), we(we_), W(W_) {
        hd = ws & 1;
        a = ws + hd;
        bb = (we - 4) & ~1;
        nw = (mode == 1 && bb >= a) ? ((bb - a + 3) >> 2) + 1 : 0;
        cend = nw ? bb + 4 : a;
    }
    __device__ __forceinline__ int count() const { return hd + nw + (we - cend); }      // mode 1
    __device__ __forceinline__ int col(int t) const {
        if (mode == 0) return max(min(ws + t, we - 1), ws);
        if (t < hd) return ws;
        const int k = t - hd;
        if (k < nw) return W + (min(a + 4 * k, bb) >> 1);
        return min(cend + (k - nw), we - 1);
    }
};

template <int CI>
__device__ __forceinline__ void half_store(const float (&m)[CI], const int (&pos)[CI], bool empty, const uint8_t* __restrict__ otab,
                                           int W, float scl, uint32_t* __restrict__ sp) {
#pragma unroll
    for (int c = 0; c < CI; ++c) {
        float outv = empty ? 0.f : -FLT_MAX;
        int idx = -1;
        if (pos[c] >= 0) {
            outv = m[c];
            const int hh = pos[c] >> 8, cc = pos[c] & 255;
            if (cc >= W) {
                idx = hh * W + ((cc - W) << 1);
                idx += otab[(idx >> 1) * CI + c];
            } else {
                idx = hh * W + cc;
            }
        }
        sp[c * kPP] = pack_out(outv, scl, idx);
    }
}

#define SOSWSOD_HALF_UPDATE(COLV)                                \
    {                                                            \
        float v[CI];                                             \
        ld_cell<CI>(v, src + (COLV) * CI);                       \
        const int code = rowcode + (COLV);                       \
        _Pragma("unroll") for (int c = 0; c < CI; ++c)           \
            if (v[c] > m[c]) {                                   \
                m[c] = v[c];                                     \
                pos[c] = code;                                   \
            }                                                    \
    }

// both steps (bin rows sub and sub + 4) of a roi with the lane's NS loads per bin row
template <int CI, int NS>
__device__ __forceinline__ void pool_roi_cols(const float* __restrict__ rows, const uint8_t* __restrict__ otab, int pitch,
                                              int W, const BinCols& bc, uint32_t hb0, uint32_t hb1, int sub, int pw,
                                              bool active, int bw, float scl, uint32_t* __restrict__ sp) {
    int col[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) col[t] = bc.col(t);
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        const int ph = sub + 4 * s;
        const bool on = active && ph < kPlanP;
        const uint32_t hb = s ? hb1 : hb0;
        const int hs = (int)(hb & 0xFFFFu), he = (int)(hb >> 16);
        const int nrows = (on && bw > 0) ? he - hs : 0;
        const int nrmax = (int)__reduce_max_sync(FULL_MASK, (unsigned)nrows);
        if (!on) continue;
        float m[CI];
        int pos[CI];
#pragma unroll
        for (int c = 0; c < CI; ++c) {
            m[c] = -FLT_MAX;
            pos[c] = -1;
        }
        const float* src = rows + hs * pitch;
        int rowcode = hs << 8;
        for (int h = 0; h < nrmax; ++h, src += pitch, rowcode += 256) {
            if (h < nrows) {
#pragma unroll
                for (int t = 0; t < NS; ++t) SOSWSOD_HALF_UPDATE(col[t])
            }
        }
        half_store<CI>(m, pos, nrows <= 0, otab, W, scl, sp + ph * kPlanP + pw);
    }
}

// any bin shape (clipped bins, bins wider than 16 cells, mixed widths): the same cover with per-lane loop bounds
template <int CI>
__device__ __forceinline__ void pool_bin_wide(const float* __restrict__ rows, const uint8_t* __restrict__ otab, int pitch,
                                              int W, int hs, int he, int ws, int we, float scl, uint32_t* __restrict__ sp) {
    float m[CI];
    int pos[CI];
#pragma unroll
    for (int c = 0; c < CI; ++c) {
        m[c] = -FLT_MAX;
        pos[c] = -1;
    }
    const bool empty = he <= hs || we <= ws;
    if (!empty) {
        const float* src = rows + hs * pitch;
        int rowcode = hs << 8;
        for (int h = hs; h < he; ++h, src += pitch, rowcode += 256) {
            int x = ws;
            if (x & 1) {
                SOSWSOD_HALF_UPDATE(x)
                ++x;
            }
            for (; x + 4 <= we; x += 4) SOSWSOD_HALF_UPDATE(W + (x >> 1))
            for (; x < we; ++x) SOSWSOD_HALF_UPDATE(x)
        }
    }
    half_store<CI>(m, pos, empty, otab, W, scl, sp);
}

template <int CI>
__global__ void __launch_bounds__(kFastThreads, 1)
roi_pool_fwd_half_kernel(const float* __restrict__ feat, int C, int H, int W, const int* __restrict__ img_start,
                         const int* __restrict__ order, const RoiRecord* __restrict__ rec,
                         uint16_t* __restrict__ argmax_u16, __nv_bfloat16* __restrict__ out_bf16, long long ld_bf16,
                         int chunks) {
    extern __shared__ __align__(16) float smem[];
    const int HW = H * W;
    // floats per row: W cells, then W/2 window entries, at an ODD pitch in cells: the maps' widths are multiples of 16 and
    // with an even pitch the lanes of one bin column in different bin rows hit the same banks on every load (958 -> 907 us
    // at 96x128; the same padding made the full-table kernels 1-3 % slower and is not applied there)
    const int pitch = ((W + (W >> 1)) | 1) * CI;
    float* rows = smem;                                 // [H][pitch]
    uint32_t* stage = reinterpret_cast<uint32_t*>(rows + (size_t)H * pitch);           // [warps][CI * 49]
    uint8_t* otab = reinterpret_cast<uint8_t*>(stage + kFastWarps * CI * kPP);         // [HW / 2][CI] offset 0..3 of the first maximum
    const int groups = C / CI;
    const int b = blockIdx.x / groups;
    const int c0 = (blockIdx.x % groups) * CI;

    {
        const float* src = feat + ((size_t)b * C + c0) * HW;
        for (int i = threadIdx.x; i < HW; i += kFastThreads) {
            const int h = i / W, x = i - h * W;
            float v[CI];
#pragma unroll
            for (int c = 0; c < CI; ++c) v[c] = __ldg(src + (size_t)c * HW + i);
            st_cell<CI>(rows + h * pitch + x * CI, v);
        }
        __syncthreads();
        for (int j = threadIdx.x; j < (HW >> 1); j += kFastThreads) {
            const int i = j << 1;
            const int h = i / W, x = i - h * W;
            float* rp = rows + h * pitch;
            float v[CI];
            uint32_t o = 0;
#pragma unroll
            for (int c = 0; c < CI; ++c) v[c] = -FLT_MAX;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float u[CI];
                ld_cell<CI>(u, rp + min(x + k, W - 1) * CI);      // a window cut by the row end is used by no bin
#pragma unroll
                for (int c = 0; c < CI; ++c)
                    if (u[c] > v[c]) {
                        v[c] = u[c];
                        o = (o & ~(0xFFu << (8 * c))) | ((uint32_t)k << (8 * c));
                    }
            }
            st_cell<CI>(rp + (W + (x >> 1)) * CI, v);
            if constexpr (CI == 4) reinterpret_cast<uint32_t*>(otab)[j] = o;
            else reinterpret_cast<uint16_t*>(otab)[j] = (uint16_t)o;
        }
        __syncthreads();
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / kPlanP, pw = lane - sub * kPlanP;
    const bool active = lane < 4 * kPlanP;
    uint32_t* sp = stage + warp * (CI * kPP);
    constexpr int n_out = CI * kPP;

    const int seg_lo = img_start[b], seg_hi = img_start[b + 1];
    const int per = (seg_hi - seg_lo + chunks - 1) / chunks;
    const int lo = seg_lo + blockIdx.y * per;
    const int hi = min(lo + per, seg_hi);
    for (int i = lo + warp; i < hi; i += kFastWarps) {
        const int r = __ldg(order + i);
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(rec + r);
        uint32_t hb0 = 0, hb1 = 0, wb = 0;
        if (active) {
            hb0 = __ldg(rw + sub);
            if (sub + 4 < kPlanP) hb1 = __ldg(rw + sub + 4);
            wb = __ldg(rw + kPlanP + pw);
        }
        const float scl = __uint_as_float(__ldg(rw + 15));
        const int ws = (int)(wb & 0xFFFFu), we = (int)(wb >> 16);
        const int bw = active ? we - ws : 0;
        const int bwmax = (int)__reduce_max_sync(FULL_MASK, (unsigned)bw);
        const int bwmin = (int)__reduce_min_sync(FULL_MASK, active ? (unsigned)bw : 0xFFFFu);
        // 0: cells one by one (bins <= 4 wide), 1: windows + edge cells with a uniform load count, 2: per-lane loops
        const int mode = bwmax <= 4 ? 0 : ((bwmax <= 16 && bwmin >= 4) ? 1 : 2);
        if (mode == 2) {
            if (active) {
                pool_bin_wide<CI>(rows, otab, pitch, W, (int)(hb0 & 0xFFFFu), (int)(hb0 >> 16), ws, we, scl, sp + sub * kPlanP + pw);
                if (sub + 4 < kPlanP)
                    pool_bin_wide<CI>(rows, otab, pitch, W, (int)(hb1 & 0xFFFFu), (int)(hb1 >> 16), ws, we, scl,
                                      sp + (sub + 4) * kPlanP + pw);
            }
        } else {
            const BinCols bc(mode, ws, we, W);
            const int nslots = mode == 0 ? max(bwmax, 1) : (int)__reduce_max_sync(FULL_MASK, (unsigned)(bw > 0 ? bc.count() : 0));
#define SOSWSOD_COLS(NSV) \
    case NSV: pool_roi_cols<CI, NSV>(rows, otab, pitch, W, bc, hb0, hb1, sub, pw, active, bw, scl, sp); break;
            switch (nslots) {
                SOSWSOD_COLS(1) SOSWSOD_COLS(2) SOSWSOD_COLS(3) SOSWSOD_COLS(4) SOSWSOD_COLS(5) SOSWSOD_COLS(6)
            }
#undef SOSWSOD_COLS
        }
        __syncwarp();
        uint16_t* adst = argmax_u16 + ((size_t)r * C + c0) * kPP;
        uint16_t* vdst = reinterpret_cast<uint16_t*>(out_bf16 + (size_t)r * ld_bf16 + (size_t)c0 * kPP);
        if (CI == 4) {
            for (int q = lane; q < n_out / 4; q += 32) {
                const uint4 w4 = *reinterpret_cast<const uint4*>(sp + 4 * q);
                uint2 av, vv;
                av.x = (w4.x & 0xFFFFu) | (w4.y << 16);
                av.y = (w4.z & 0xFFFFu) | (w4.w << 16);
                vv.x = (w4.x >> 16) | (w4.y & 0xFFFF0000u);
                vv.y = (w4.z >> 16) | (w4.w & 0xFFFF0000u);
                reinterpret_cast<uint2*>(adst)[q] = av;
                reinterpret_cast<uint2*>(vdst)[q] = vv;
            }
        } else {
            for (int q = lane; q < n_out / 2; q += 32) {
                const uint2 w2 = *reinterpret_cast<const uint2*>(sp + 2 * q);
                reinterpret_cast<uint32_t*>(adst)[q] = (w2.x & 0xFFFFu) | (w2.y << 16);
                reinterpret_cast<uint32_t*>(vdst)[q] = (w2.x >> 16) | (w2.y & 0xFFFF0000u);
            }
        }
        __syncwarp();
    }
}
#undef SOSWSOD_HALF_UPDATE

static int fwd_fast_chunks(int groups, int n, int R) {
    const int sms = device_num_sms();
    const int per_img = max(1, R / max(n, 1));
    const int max_chunks = max(1, min(16, per_img / (2 * kFastWarps)));
    int chunks = 1;
    double best = 1e30;
    for (int ch = 1; ch <= max_chunks; ++ch) {
        const int waves = (groups * ch + sms - 1) / sms;
        const double cost = (double)waves / ch + 0.004 * ch;
        if (cost < best - 1e-9) {
            best = cost;
            chunks = ch;
        }
    }
    return chunks;
}

template <int CI>
static int launch_fwd_half_ci(const float* feat, int n, int c, int h, int w, int R, const PlanView& pv, uint16_t* a16,
                              __nv_bfloat16* obf, long long ld, size_t smem, cudaStream_t st) {
    const int groups = n * (c / CI);
    const int chunks = fwd_fast_chunks(groups, n, R);
    auto kern = roi_pool_fwd_half_kernel<CI>;
    SOSWSOD_ENSURE_SMEM(kern, smem);
    kern<<<dim3(groups, chunks), kFastThreads, smem, st>>>(feat, c, h, w, pv.img_start, pv.order, pv.rec, a16, obf, ld, chunks);
    SOSWSOD_CHECK_LAUNCH();
    return 1;
}

template <int CI>
static int launch_fwd_fast_ci(const float* feat, int n, int c, int h, int w, int R, const PlanView& pv, uint16_t* a16,
                              __nv_bfloat16* obf, long long ld, int cells_pad, size_t smem, cudaStream_t st) {
    const int groups = n * (c / CI);
    const int sms = device_num_sms();
    // rois of an image are split over `chunks` CTAs per channel group: fill whole waves of the SMs (1 CTA/SM) with
    // the smallest tail, but keep chunks long enough that staging the planes stays a small part of a CTA's work
    const int per_img = max(1, R / max(n, 1));
    const int max_chunks = max(1, min(16, per_img / (2 * kFastWarps)));
    int chunks = 1;
    double best = 1e30;
    for (int ch = 1; ch <= max_chunks; ++ch) {
        const int waves = (groups * ch + sms - 1) / sms;
        const double cost = (double)waves / ch + 0.004 * ch;
        if (cost < best - 1e-9) {
            best = cost;
            chunks = ch;
        }
    }
    auto kern = roi_pool_fwd_fast_kernel<CI>;
    SOSWSOD_ENSURE_SMEM(kern, smem);
    kern<<<dim3(groups, chunks), kFastThreads, smem, st>>>(feat, c, h, w, pv.img_start, pv.order, pv.rec, a16, obf, ld,
                                                          cells_pad, chunks);
    SOSWSOD_CHECK_LAUNCH();
    return 1;
}

int launch_fwd_fast(const float* feat, int n, int c, int h, int w, int R, const void* plan, uint16_t* argmax_u16,
                    __nv_bfloat16* out_bf16, long long ld_bf16, cudaStream_t st) {
    if (!argmax_u16 || !out_bf16 || n > kPlanMaxImages || (long long)h * w >= 65535) return 0;
    const int HW = h * w;
    const int cells_pad = (HW + 4 + 3) / 4 * 4;
    const size_t max_smem = (size_t)device_max_smem();
    const PlanView pv = plan_view(plan, R);
    const bool al8 = (reinterpret_cast<uintptr_t>(argmax_u16) & 7) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0 &&
                     ((ld_bf16 * 2) & 7) == 0 && (((long long)c * kPP * 2) & 7) == 0;
    const bool al4 = (reinterpret_cast<uintptr_t>(argmax_u16) & 3) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 3) == 0 &&
                     ((ld_bf16 * 2) & 3) == 0 && (((long long)c * kPP * 2) & 3) == 0;
    // four channels per CTA before two; the half table (two channels) only where the full one does not fit
    auto full_smem = [&](int ci) { return (size_t)cells_pad * ci * 4 * 2 + (size_t)kFastWarps * ci * kPP * 4 + (size_t)cells_pad * ci; };
    auto half_smem = [&](int ci) { return (size_t)h * ((w + w / 2) | 1) * ci * 4 + (size_t)kFastWarps * ci * kPP * 4 + (size_t)(HW / 2) * ci; };
    const bool half_ok = (w % 2 == 0) && w + w / 2 <= 256 && h <= 255;
    if (c % 4 == 0 && al8) {
        if (full_smem(4) <= max_smem)
            return launch_fwd_fast_ci<4>(feat, n, c, h, w, R, pv, argmax_u16, out_bf16, ld_bf16, cells_pad, full_smem(4), st);
    }
    if (c % 2 == 0 && al4) {
        if (full_smem(2) <= max_smem)
            return launch_fwd_fast_ci<2>(feat, n, c, h, w, R, pv, argmax_u16, out_bf16, ld_bf16, cells_pad, full_smem(2), st);
        if (half_ok && half_smem(2) <= max_smem)
            return launch_fwd_half_ci<2>(feat, n, c, h, w, R, pv, argmax_u16, out_bf16, ld_bf16, half_smem(2), st);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// A CTA owns the gradient planes of CT consecutive channels of one image (optionally one band of rows of them) in
// shared memory; ONE accumulator warp is the only writer of a pair of planes (lanes 0-15 the first, lanes 16-31 the
// second), so there is no atomic and no cross-warp race, and the accumulation order is fixed (deterministic).
//
// Conflict freedom inside a warp step comes from PROPOSAL-BIN OWNERSHIP: the arg-max of a bin lies inside the bin,
// and two bins of one roi can only share a cell when their row ranges AND column ranges overlap.  The plan holds,
// per roi, the smallest strides (mh, mw) such that bins mh rows (mw columns) apart are disjoint (2 x 2 for every roi
// at least 7 cells high and wide, larger for tiny rois whose bins repeat cells).  A 16-lane half-warp = 4 x 4 blocks
// of mh x mw bins; a step takes ONE colour (i, j) -- bin (la*mh + i, lb*mw + j) of every block -- whose bins are
// pairwise disjoint, so every lane does a plain read-add-write on its own cell.
constexpr int kBwdFastMaxCT = 8;
constexpr int kBwdFastMaxRT = 16;

struct BwdFastCfg {
    int CT, bands, band_rows, nbox, BW, stages, RT, plane_stride;
    size_t smem;
};

// ------------------------------------------------------------------------------------------------
// backward, queued.  The ordered read-add-write chain of a plane pair is the critical path (~200 cycles per roi);
// everything around it (operand fetch, address arithmetic) is taken off that path: every plane pair has ONE
// accumulator warp that does nothing but consume a stream of ready-made steps, fed through a shared-memory queue.
// (Measured alternatives, see profiles/README.md: warps handing a turn token around, 628 us; merging a roi's
// duplicate cells in the prep warps so that its steps become independent, 575 us -- the extra shuffles make the prep
// warps the bottleneck; this kernel, 440-525 us.)
//
//   TMA warp      streams [RT rois x BW columns] tiles of arg-max / grad_out through the mbarrier ring and, per roi,
//                 publishes (scale, colour strides, first chunk number): the rois' colour steps are cut into CHUNKS
//                 of 4 steps, numbered consecutively over the image's rois (a warp scan per tile), so that the
//                 accumulator sees one uniform stream.  Plan records are fetched 32 rois at a time, one window ahead.
//   prep warps    P per plane pair, rois dealt round-robin.  A prep warp turns each chunk of its roi into 4 x 32
//                 (shared address, scaled gradient) pairs and stores them into queue slot (chunk number mod Q) once
//                 the accumulator has consumed the slot's previous occupant.
//   accumulator   chunk by chunk: entries of the next chunk are fetched while the current one accumulates.
//
// Queue protocol, no fences and no mbarriers on the accumulator's path: an entry is ONE 64-bit shared store
// {address (18 bits) | tag << 19, value}; the tag names the slot's use (chunk number / Q).  A prep warp writes the 4
// steps of a chunk in order 0..3; the accumulator reads them in order 3..0 and checks step 3's tag in every lane:
// shared-memory requests of an SM are served in order, so a current step 3 implies current steps 0..2.  Slots return
// to the prep warps through one counter per plane pair (chunks consumed so far; single writer).
// Result: deterministic (fixed order per plane), atomic-free.
constexpr int kBwdQSteps = 4;                       // colour steps per chunk / queue slot
constexpr int kBwdQSlotBytes = kBwdQSteps * 32 * 8;
constexpr int kBwdQMaxP = 6;
constexpr int kBwdQMaxQ = 8;

struct BwdQCfg {
    BwdFastCfg b;
    int P;             // prep warps per plane pair
    int LQ;            // log2 of the queue depth Q (slots per plane pair)
};

struct __align__(16) BwdMetaQ {
    float scale;
    int code;    // 0 = roi of another image (no chunks); else 1 | mh << 8 | mw << 16
    int cbase;   // number of the roi's first chunk in the image's stream
    int nch;     // its number of chunks
};

__device__ __forceinline__ int bwdq_chunks_of(uint32_t tail_x, int b) {
    // tail_x = bytes 56..59 of a RoiRecord: mh | mw << 8 | batch << 16
    if ((int)(tail_x >> 16) != b) return 0;
    const int nsteps = (int)(tail_x & 0xFFu) * (int)((tail_x >> 8) & 0xFFu);
    return (nsteps + kBwdQSteps - 1) / kBwdQSteps;
}

// steps 3, 2, 1, 0 in that order (see the protocol above)
__device__ __forceinline__ void q_load(uint32_t qa, uint32_t (&lo)[kBwdQSteps], uint32_t (&hi)[kBwdQSteps]) {
#pragma unroll
    for (int k = kBwdQSteps - 1; k >= 0; --k)
        asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo[k]), "=r"(hi[k]) : "r"(qa + 256u * k) : "memory");
}

template <typename GradT>
__global__ void __launch_bounds__((kBwdFastMaxCT / 2 * (1 + kBwdQMaxP) + 1) * 32, 1)
roi_pool_bwd_q_kernel(const __grid_constant__ CUtensorMap tmap_arg, const __grid_constant__ CUtensorMap tmap_grad,
                      const int* __restrict__ img_start, const int* __restrict__ order,
                      const RoiRecord* __restrict__ rec, int C, int H, int W, float* __restrict__ grad_feat, BwdQCfg qc) {
    constexpr int PP = kPP;
    extern __shared__ uint8_t smem_raw[];
    const BwdFastCfg& cfg = qc.b;
    const int CT = cfg.CT, BW = cfg.BW, nbox = cfg.nbox, S = cfg.stages, RT = cfg.RT;
    const int P = qc.P, LQ = qc.LQ, Q = 1 << LQ;
    const int NP = (CT + 1) >> 1;   // plane pairs = accumulator warps
    const int NPREP = NP * P;
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    float* planes = reinterpret_cast<float*>(gen_base);                          // [CT][plane_stride]
    const uint32_t plane_bytes = (uint32_t)CT * cfg.plane_stride * 4;            // multiple of 128
    const uint32_t arg_box = (uint32_t)RT * BW * 2;                              // multiples of 128 (checked by the host)
    const uint32_t grad_box = (uint32_t)RT * BW * sizeof(GradT);
    const uint32_t arg_stage = nbox * arg_box, grad_stage = nbox * grad_box;
    const uint32_t ring_off = plane_bytes;
    const uint32_t queue_off = ring_off + S * (arg_stage + grad_stage);          // [NP][Q][4][32] x 8 bytes
    const uint32_t bar_off = queue_off + (uint32_t)NP * Q * kBwdQSlotBytes;
    auto full_bar = [&](int st) { return base + bar_off + 8u * st; };
    auto empty_bar = [&](int st) { return base + bar_off + 8u * (S + st); };
    const uint32_t cons_off = bar_off + 16u * S;                                 // [NP] chunks consumed, [NP] = total
    const uint32_t meta_off = cons_off + 4u * (kBwdFastMaxCT / 2 + 4);           // 16-byte aligned
    BwdMetaQ* s_meta = reinterpret_cast<BwdMetaQ*>(gen_base + meta_off);         // [S][kBwdFastMaxRT]
    float* s_dummy = reinterpret_cast<float*>(s_meta + S * kBwdFastMaxRT);        // [32] idle-lane targets
    int* s_total = reinterpret_cast<int*>(gen_base + cons_off) + kBwdFastMaxCT / 2;

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
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(gen_base + queue_off);
        const int nz = (int)((meta_off - queue_off) / 4);      // queue entries (tag 0 = invalid), barriers, counters
        for (int i = threadIdx.x; i < nz; i += blockDim.x) z[i] = 0u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int st = 0; st < S; ++st) {
            mbar_init(full_bar(st), 1);
            mbar_init(empty_bar(st), NPREP);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // rows of this image: the plan's order is stable, so its first / last entries are the smallest / largest row
    const int seg_lo = img_start[b], seg_hi = img_start[b + 1];
    const int r_lo = seg_hi > seg_lo ? __ldg(order + seg_lo) : 0;
    const int r_hi = seg_hi > seg_lo ? __ldg(order + seg_hi - 1) + 1 : 0;
    const int nrois = r_hi - r_lo;
    const int ntiles = (nrois + RT - 1) / RT;
    {
        // length of the image's chunk stream (the accumulators' trip count)
        int mine = 0;
        for (int i = threadIdx.x; i < nrois; i += blockDim.x)
            mine += bwdq_chunks_of(__ldg(reinterpret_cast<const uint32_t*>(rec + r_lo + i) + 14), b);
        mine = warp_sum_int(mine);
        if (lane == 0 && mine) atomicAdd(s_total, mine);   // integer bookkeeping
    }
    __syncthreads();
    const int total_chunks = *s_total;
    const int total_cols = C * PP;
    const int col0 = c0 * PP;
    const int col_a = col0 & ~7;
    const int col_g = col0 & ~(16 / (int)sizeof(GradT) - 1);

    if (warp == NP + NPREP) {
        // ---- TMA warp ----
        uint32_t tx = 0;
        for (int bx = 0; bx < nbox; ++bx) {
            if (col_a + bx * BW < total_cols) tx += arg_box;
            if (col_g + bx * BW < total_cols) tx += grad_box;
        }
        // plan records (bytes 56..63) of a window of 32 rois per lane, the next window already in flight
        auto fetch = [&](int w0) {
            uint2 v = make_uint2(0xFFFF0000u, 0u);      // batch 0xFFFF: no image
            if (w0 + lane < nrois) v = __ldg(reinterpret_cast<const uint2*>(rec + r_lo + w0 + lane) + 7);
            return v;
        };
        int win = 0;
        uint2 cur = fetch(0), nxt = fetch(32);
        int running = 0;
        int st = 0;
        uint32_t phase = 0;
        for (int t = 0; t < ntiles; ++t) {
            const int r0 = t * RT;
            if (r0 >= win + 32) {
                win += 32;
                cur = nxt;
                nxt = fetch(win + 32);
            }
            const int src = r0 - win + lane;               // RT divides 32: a tile never straddles two windows
            const uint32_t tail_x = __shfl_sync(FULL_MASK, cur.x, src & 31);
            const uint32_t tail_y = __shfl_sync(FULL_MASK, cur.y, src & 31);
            BwdMetaQ meta;
            meta.scale = __uint_as_float(tail_y);
            meta.nch = (lane < RT && r0 + lane < nrois) ? bwdq_chunks_of(tail_x, b) : 0;
            meta.code = meta.nch ? (int)(1u | ((tail_x & 0xFFu) << 8) | (((tail_x >> 8) & 0xFFu) << 16)) : 0;
            int incl = meta.nch;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o) incl += up;
            }
            meta.cbase = running + incl - meta.nch;
            running += __shfl_sync(FULL_MASK, incl, 31);
            mbar_wait(empty_bar(st), phase ^ 1u);
            if (lane < RT) s_meta[st * kBwdFastMaxRT + lane] = meta;
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(full_bar(st), tx);
                const uint32_t sa = base + ring_off + st * (arg_stage + grad_stage);
                const uint32_t sg = sa + arg_stage;
                for (int bx = 0; bx < nbox; ++bx) {
                    if (col_a + bx * BW < total_cols) tma_load_2d(sa + bx * arg_box, &tmap_arg, full_bar(st), col_a + bx * BW, r_lo + r0);
                    if (col_g + bx * BW < total_cols) tma_load_2d(sg + bx * grad_box, &tmap_grad, full_bar(st), col_g + bx * BW, r_lo + r0);
                }
            }
            if (++st == S) {
                st = 0;
                phase ^= 1u;
            }
        }
    } else if (warp < NP) {
        // ---- accumulator warp of plane pair `warp` ----
        const uint32_t qbase = base + queue_off + (uint32_t)warp * Q * kBwdQSlotBytes + 8u * lane;
        const uint32_t cons_s = base + cons_off + 4u * warp;
        const uint32_t qmask = (uint32_t)Q - 1u;
        // one chunk: wait until it is complete, give its slot back, start fetching chunk seq + 1 into (nlo, nhi), then
        // run the ordered steps
        // one chunk: (wait until it is complete,) give its slot back, start fetching chunk seq + 1 into (nlo, nhi), run the
        // ordered steps; the next chunk's tag check rides in the shadow of the chain's shared-memory latencies
        auto consume = [&](uint32_t seq, bool valid, uint32_t (&lo)[kBwdQSteps], uint32_t (&hi)[kBwdQSteps],
                           uint32_t (&nlo)[kBwdQSteps], uint32_t (&nhi)[kBwdQSteps]) -> bool {
            if (!valid) {
                const uint32_t tag = 0x1000u | ((seq >> LQ) & 0xFFFu);
                do {
                    q_load(qbase + (seq & qmask) * kBwdQSlotBytes, lo, hi);
                } while (!__all_sync(FULL_MASK, (lo[kBwdQSteps - 1] >> 19) == tag));
            }
            if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(cons_s), "r"(seq + 1u) : "memory");
            const bool have_next = seq + 1u < (uint32_t)total_chunks;
            if (have_next) q_load(qbase + ((seq + 1u) & qmask) * kBwdQSlotBytes, nlo, nhi);
            const uint32_t ntag = 0x1000u | (((seq + 1u) >> LQ) & 0xFFFu);
            bool nvalid = false;
#pragma unroll
            for (int i = 0; i < kBwdQSteps; ++i) {
                const uint32_t ad = lo[i] & 0x3FFFFu;
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(ad) : "memory");
                if (i == 1) nvalid = __all_sync(FULL_MASK, have_next && (nlo[kBwdQSteps - 1] >> 19) == ntag);
                v += __uint_as_float(hi[i]);
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(ad), "f"(v) : "memory");
                __syncwarp();   // colour classes of one roi may share cells: order the steps
            }
            return nvalid;
        };
        uint32_t alo[kBwdQSteps], ahi[kBwdQSteps], blo[kBwdQSteps], bhi[kBwdQSteps];
        bool valid = false;
        for (uint32_t seq = 0; seq < (uint32_t)total_chunks; seq += 2) {
            valid = consume(seq, valid, alo, ahi, blo, bhi);
            if (seq + 1u < (uint32_t)total_chunks) valid = consume(seq + 1u, valid, blo, bhi, alo, ahi);
        }
    } else {
        // ---- preparation warp j of plane pair `pair` ----
        const int pw_ = warp - NP;
        const int pair = pw_ / P, j = pw_ - pair * P;
        const int half = lane >> 4, la = (lane >> 2) & 3, lb = lane & 3;
        const int chan = 2 * pair + half;
        const bool chan_ok = chan < CT && (c0 + chan) < C;
        const uint32_t my_s = smem_u32(planes + (chan_ok ? chan : 0) * cfg.plane_stride);
        const uint32_t dummy_s = smem_u32(s_dummy) + 4u * lane;
        const int ea0 = chan * PP + (col0 - col_a);
        const int eg0 = chan * PP + (col0 - col_g);
        const unsigned band_cells = (unsigned)(band_hi - band_lo);
        const int wrap = (RT - 1) * BW;          // elements skipped when an entry falls into the second box
        const uint32_t ring_s = base + ring_off;
        const uint32_t stage_bytes = arg_stage + grad_stage;
        const uint32_t qpair = base + queue_off + (uint32_t)pair * Q * kBwdQSlotBytes + 8u * lane;
        const uint32_t cons_s = base + cons_off + 4u * pair;
        const uint32_t qmask = (uint32_t)Q - 1u;
        constexpr int kCode22 = 1 | (2 << 8) | (2 << 16);
        int e22a[4], e22g[4];
        bool v22[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ph = 2 * la + (k >> 1), pw = 2 * lb + (k & 1);
            v22[k] = chan_ok && ph < kPlanP && pw < kPlanP;
            int e = ea0 + ph * kPlanP + pw;
            e22a[k] = e + (e >= BW ? wrap : 0);
            e = eg0 + ph * kPlanP + pw;
            e22g[k] = e + (e >= BW ? wrap : 0);
        }
        int st = 0;
        uint32_t phase = 0;
        int r = j;
        for (int t = 0; t < ntiles; ++t) {
            const int tile_end = min((t + 1) * RT, nrois);
            // every prep warp waits for every tile, also one that holds none of its rois: its arrival on the empty
            // barrier must not run ahead into the stage's previous phase
            mbar_wait(full_bar(st), phase);
            const uint32_t sa = ring_s + st * stage_bytes;      // arg-max boxes of this stage
            const uint32_t sg = sa + arg_stage;                 // grad boxes
            const BwdMetaQ* metas = s_meta + st * kBwdFastMaxRT;
            for (; r < tile_end; r += P) {
                const int rr = r - t * RT;
                const BwdMetaQ m = metas[rr];
                if (m.nch == 0) continue;                       // roi of another image
                const int mh = (m.code >> 8) & 0xFF, mw = (m.code >> 16) & 0xFF;
                const int nsteps = mh * mw;
                const int rowe = rr * BW;
                auto operand = [&](int ea, int eg, uint32_t& ad, uint32_t& vl) {
                    unsigned a, graw;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(a) : "r"(sa + 2u * (unsigned)(ea + rowe)));
                    if (sizeof(GradT) == 2) {
                        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(graw) : "r"(sg + 2u * (unsigned)(eg + rowe)));
                        graw <<= 16;
                    } else {
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(graw) : "r"(sg + 4u * (unsigned)(eg + rowe)));
                    }
                    const unsigned rel = a - (unsigned)band_lo;   // empty bin (0xFFFF) and other bands fail the test
                    const bool ok = rel < band_cells;
                    ad = ok ? my_s + 4u * rel : dummy_s;
                    vl = ok ? __float_as_uint(__uint_as_float(graw) * m.scale) : 0u;
                };
                auto publish = [&](const uint32_t (&ad)[kBwdQSteps], const uint32_t (&vl)[kBwdQSteps], uint32_t seq) {
                    uint32_t consumed;
                    do {   // the slot's previous occupant is chunk seq - Q
                        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(consumed) : "r"(cons_s) : "memory");
                    } while ((int)(seq - consumed) >= Q);
                    const uint32_t hdr = (0x1000u | ((seq >> LQ) & 0xFFFu)) << 19;
                    const uint32_t qa = qpair + (seq & qmask) * kBwdQSlotBytes;
#pragma unroll
                    for (int i = 0; i < kBwdQSteps; ++i)
                        asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(qa + 256u * i), "r"(ad[i] | hdr), "r"(vl[i]) : "memory");
                };
                uint32_t ad[kBwdQSteps], vl[kBwdQSteps];
                if (m.code == kCode22) {
                    // strides 2 x 2 (every roi at least 7 x 7 cells): one chunk, the lane's four bins at constant offsets
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        ad[k] = dummy_s;
                        vl[k] = 0u;
                        if (v22[k]) operand(e22a[k], e22g[k], ad[k], vl[k]);
                    }
                    publish(ad, vl, (uint32_t)m.cbase);
                } else {
                    const int bin0 = la * mh * kPlanP + lb * mw;
                    const int ih = chan_ok ? min(mh, kPlanP - la * mh) : 0;   // <= 0: block outside the grid
                    const int jw = min(mw, kPlanP - lb * mw);
                    int i = 0, jj = 0;
                    uint32_t seq = (uint32_t)m.cbase;
                    for (int s0 = 0; s0 < nsteps; s0 += kBwdQSteps, ++seq) {
                        const int n = min(kBwdQSteps, nsteps - s0);
#pragma unroll
                        for (int k = 0; k < kBwdQSteps; ++k) {
                            ad[k] = dummy_s;
                            vl[k] = 0u;
                            if (k < n) {
                                if (i < ih && jj < jw) {
                                    const int bin = bin0 + i * kPlanP + jj;
                                    int ea = ea0 + bin, eg = eg0 + bin;
                                    ea += (ea >= BW ? wrap : 0);
                                    eg += (eg >= BW ? wrap : 0);
                                    operand(ea, eg, ad[k], vl[k]);
                                }
                                if (++jj == mw) {
                                    jj = 0;
                                    ++i;
                                }
                            }
                        }
                        publish(ad, vl, seq);
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

// Tuning overrides, read ONCE per process (0 = let the search decide).
static int env_int_once(const char* name, int* cache) {
    if (*cache < 0) {
        const char* e = getenv(name);
        *cache = e ? atoi(e) : 0;
    }
    return *cache;
}

static bool pick_bwd_q_cfg(int n, int c, int h, int w, int grad_bytes, BwdQCfg* out) {
    const int max_smem = device_max_smem();
    const int sms = device_num_sms();
    const int align_elems = 16 / (2 < grad_bytes ? 2 : grad_bytes);
    // tuning overrides (0 = let the search decide)
    static int env_p = -1, env_lq = -1;
    const int force_p = env_int_once("SOSWSOD_BWDQ_P", &env_p), force_lq = env_int_once("SOSWSOD_BWDQ_LQ", &env_lq);
    bool found = false;
    double best_cost = 0;
    for (int bands = 1; bands <= 64; ++bands) {
        const int band_rows = (h + bands - 1) / bands;
        if (bands > 1 && (long long)(bands - 1) * band_rows >= h) continue;  // empty last band
        const int plane_stride = ((band_rows * w + 31) / 32) * 32;
        for (int CT = kBwdFastMaxCT; CT >= 1; --CT) {
            if (CT > c) continue;
            const int NP = (CT + 1) / 2;
            const int cols = CT * kPP + align_elems - 1;   // + the alignment remainder of the first column
            const int nbox = (cols + 247) / 248;
            const int BW = (((cols + nbox - 1) / nbox) + 7) / 8 * 8;
            if (BW > 256 || nbox > 2) continue;
            for (int P = kBwdQMaxP; P >= 2; --P)
                for (int LQ = 3; LQ >= 1; --LQ) {
                    if ((force_p && P != force_p) || (force_lq && LQ != force_lq)) continue;
                    const size_t fixed = (size_t)CT * plane_stride * 4 + (size_t)NP * (1 << LQ) * kBwdQSlotBytes + 128 /*align*/ +
                                         2304 /*barriers, counters, roi meta, dummies*/;
                    for (int RT = kBwdFastMaxRT; RT >= 4; RT >>= 1) {
                        if (((size_t)RT * BW * 2) % 128 != 0) continue;   // every TMA box starts 128-byte aligned
                        const size_t stage = (size_t)nbox * RT * BW * (2 + grad_bytes);
                        if (fixed + 2 * stage > (size_t)max_smem) continue;
                        int stages = (int)(((size_t)max_smem - fixed) / stage);
                        if (stages > 6) stages = 6;
                        const long long ctas = (long long)n * ((c + CT - 1) / CT) * bands;
                        const long long waves = (ctas + sms - 1) / sms;
                        // time ~ waves (every CTA streams all rois of its image; its chains advance together); then
                        // enough prep warps and queue slots to keep the accumulators fed, then a ring of >= 3 stages
                        const double cost = (double)waves + 0.02 * abs(4 - P) + 0.01 * (3 - LQ) +
                                            (stages < 3 ? 0.015 : 0.0) + ((CT & 1) ? 0.01 : 0.0);
                        if (!found || cost < best_cost - 1e-9) {
                            found = true;
                            best_cost = cost;
                            out->b.CT = CT;
                            out->b.bands = bands;
                            out->b.band_rows = band_rows;
                            out->b.nbox = nbox;
                            out->b.BW = BW;
                            out->b.stages = stages;
                            out->b.RT = RT;
                            out->b.plane_stride = plane_stride;
                            out->b.smem = fixed + (size_t)stages * stage;
                            out->P = P;
                            out->LQ = LQ;
                        }
                        if (stages >= 3) break;
                    }
                }
        }
        if (found && best_cost < 1.5) break;
    }
    return found;
}

template <typename GradT>
static int launch_bwd_q_t(const void* grad, long long ld_grad, const uint16_t* argmax, int R, const void* plan, int n, int c,
                          int h, int w, float* grad_feat, cudaStream_t st) {
    // the configuration search depends on the shape only: remembered per host thread for the shapes of a step
    struct Memo {
        int n, c, h, w, found;
        BwdQCfg qc;
    };
    static thread_local Memo memo[4] = {};
    static thread_local int memo_next = 0;
    const Memo* hit = nullptr;
    for (const Memo& m : memo)
        if (m.found != 0 && m.n == n && m.c == c && m.h == h && m.w == w) hit = &m;
    if (!hit) {
        Memo& m = memo[memo_next];
        memo_next = (memo_next + 1) & 3;
        m.n = n; m.c = c; m.h = h; m.w = w;
        m.found = pick_bwd_q_cfg(n, c, h, w, (int)sizeof(GradT), &m.qc) ? 1 : -1;
        hit = &m;
    }
    if (hit->found < 0) return 0;
    const BwdQCfg qc = hit->qc;
    const BwdFastCfg& cfg = qc.b;
    CUtensorMap ta, tg;
    int rc = make_tmap_2d(&ta, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, argmax, R, (long long)c * kPP, (long long)c * kPP, cfg.BW,
                          cfg.RT, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    rc = make_tmap_2d(&tg, sizeof(GradT) == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                      (int)sizeof(GradT), grad, R, (long long)c * kPP, ld_grad, cfg.BW, cfg.RT, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    const PlanView pv = plan_view(plan, R);
    const int grid = n * ((c + cfg.CT - 1) / cfg.CT) * cfg.bands;
    const int threads = ((cfg.CT + 1) / 2 * (1 + qc.P) + 1) * 32;
    auto kern = roi_pool_bwd_q_kernel<GradT>;
    SOSWSOD_ENSURE_SMEM(kern, cfg.smem);
    kern<<<grid, threads, cfg.smem, st>>>(ta, tg, pv.img_start, pv.order, pv.rec, c, h, w, grad_feat, qc);
    SOSWSOD_CHECK_LAUNCH();
    return 1;
}

// Returns 1 when the queued kernel ran, 0 when no configuration fits (the caller takes the general kernel), < 0 on error.
int launch_bwd_fast(const void* grad, int grad_dtype, long long ld_grad, const uint16_t* argmax, int R, const void* plan,
                    int n, int c, int h, int w, float* grad_feat, cudaStream_t st) {
    if (n > kPlanMaxImages) return 0;
    const int gb = grad_dtype == SOSWSOD_DTYPE_BF16 ? 2 : 4;
    const bool aligned = ((uintptr_t)grad & 15) == 0 && ((uintptr_t)argmax & 15) == 0 && ((ld_grad * gb) & 15) == 0 &&
                         (((long long)c * kPP * 2) & 15) == 0;
    if (!aligned) return 0;
    return grad_dtype == SOSWSOD_DTYPE_BF16
               ? launch_bwd_q_t<__nv_bfloat16>(grad, ld_grad, argmax, R, plan, n, c, h, w, grad_feat, st)
               : launch_bwd_q_t<float>(grad, ld_grad, argmax, R, plan, n, c, h, w, grad_feat, st);
}

}  // namespace soswsod
