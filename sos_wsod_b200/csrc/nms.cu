// Test-time post-processing (kernel (5) of the hot path):
//   predict         -- mean over the K refinement branches of softmax(logits_k) and of the box deltas, then
//                      Box2BoxTransform.apply_deltas                       (fast_rcnn_oicr.py:674-735)
//   tta_accumulate  -- inverse view transform + running mean over views    (test_time_augmentation_avg.py:349-371)
//   nms             -- greedy NMS with the torchvision semantics the reference calls through
//                      uwsod/detectron2/layers/nms.py:6-29: score-descending (ties: lower index first),
//                      suppress when IoU > thr (strict), IoU = inter / (a + b - inter) in IEEE fp32
//   detect          -- fast_rcnn_inference_single_image (fast_rcnn_oicr.py:86-148): finite-row filter, drop the
//                      background column, clip, score > thr, per-class NMS, global top-k
//
// NMS is the 64-box bitmask algorithm (the in-repo restatement is
// uwsod/detectron2/layers/csrc/nms_rotated/nms_rotated_cuda.cu:21-143) with two changes: the sort is an
// in-shared-memory bitonic sort of (score, index) keys inside the same pipeline, and the serial reduce over
// the mask runs on the device (one CTA per class, 64 boxes resolved per step from registers) instead of a
// D2H copy plus a host loop -- no synchronisation with the host anywhere.
#include "common.cuh"

namespace soswsod {

constexpr int kNmsMaxN = 16384;
constexpr int kSortThreads = 1024;

// list l of `num_lists`; element i of n.  mode 0: plain nms (scores[i], boxes[i]); mode 1: detect.
struct NmsSource {
    int mode;
    const float* boxes;       // mode 0: [n,4]; mode 1: pred_boxes [R, 4C]
    const float* scores;      // mode 0: [n];   mode 1: probs [R, C+1]
    const uint8_t* rowvalid;  // mode 1
    int C;
    float img_h, img_w, score_thr;
};

__device__ __forceinline__ float4 clip_box(float4 b, float w, float h) {
    // Boxes.clip (detectron2/structures/boxes.py:183-196): clamp(min=0, max=w|h)
    b.x = fminf(fmaxf(b.x, 0.f), w);
    b.y = fminf(fmaxf(b.y, 0.f), h);
    b.z = fminf(fmaxf(b.z, 0.f), w);
    b.w = fminf(fmaxf(b.w, 0.f), h);
    return b;
}

__device__ __forceinline__ bool nms_fetch(const NmsSource& s, int l, int i, float& score, float4& box) {
    if (s.mode == 0) {
        score = s.scores[i];
        box = reinterpret_cast<const float4*>(s.boxes)[i];
        return true;
    }
    if (!s.rowvalid[i]) return false;
    score = s.scores[(size_t)i * (s.C + 1) + l];
    if (!(score > s.score_thr)) return false;
    box = clip_box(reinterpret_cast<const float4*>(s.boxes)[(size_t)i * s.C + l], s.img_w, s.img_h);
    return true;
}

// grid = num_lists.  Sorts the valid entries of list l by (score desc, index asc); writes the sorted original
// indices, the sorted (clipped) boxes and the count.
__global__ void __launch_bounds__(kSortThreads)
nms_sort_kernel(NmsSource src, int n, int n_pow2, int32_t* __restrict__ sorted_idx, float4* __restrict__ sorted_box,
                int32_t* __restrict__ counts) {
    extern __shared__ __align__(16) unsigned long long keys[];
    __shared__ int s_count;
    const int l = blockIdx.x;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        unsigned long long key = 0ull;
        if (i < n) {
            float sc;
            float4 bx;
            if (nms_fetch(src, l, i, sc, bx)) {
                // +1 keeps a valid key non-zero even for the lowest representable score
                key = ((unsigned long long)float_to_ordered(sc) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
                if (key == 0ull) key = 1ull;
                ++local;
            }
        }
        keys[i] = key;
    }
    local = warp_sum_int(local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_count, local);
    __syncthreads();
    bitonic_sort_desc_u64(keys, n_pow2);
    const int cnt = s_count;
    if (threadIdx.x == 0) counts[l] = cnt;
    for (int p = threadIdx.x; p < cnt; p += blockDim.x) {
        const int i = (int)(0xFFFFFFFFu - (uint32_t)(keys[p] & 0xFFFFFFFFull));
        float sc;
        float4 bx;
        nms_fetch(src, l, i, sc, bx);
        sorted_idx[(size_t)l * n + p] = i;
        sorted_box[(size_t)l * n + p] = bx;
    }
}

// grid (nb, nb, num_lists), block 64.  mask[l][i][jb] bit t set <=> j = jb*64+t > i and IoU(i, j) > thr.
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4* __restrict__ sorted_box, const int32_t* __restrict__ counts, int n, int nb, float thr,
                unsigned long long* __restrict__ mask) {
    const int l = blockIdx.z;
    const int cnt = counts[l];
    const int ib = blockIdx.y, jb = blockIdx.x;
    if (jb < ib || ib * 64 >= cnt || jb * 64 >= cnt) return;
    __shared__ float4 cb[64];
    const float4* bx = sorted_box + (size_t)l * n;
    const int jn = min(cnt - jb * 64, 64);
    if ((int)threadIdx.x < jn) cb[threadIdx.x] = bx[jb * 64 + threadIdx.x];
    __syncthreads();
    const int i = ib * 64 + threadIdx.x;
    if (i >= cnt) return;
    const float4 a = bx[i];
    unsigned long long bits = 0ull;
    const int start = (ib == jb) ? threadIdx.x + 1 : 0;
    for (int t = start; t < jn; ++t)
        if (box_iou_nms_rn(a, cb[t]) > thr) bits |= 1ull << t;
    mask[((size_t)l * n + i) * nb + jb] = bits;
}

// grid num_lists, block 256.  Serial-in-blocks reduce of the mask; emits the kept sorted positions (at most
// `max_keep` of them) in order.  kept_pos [num_lists, max_keep]; kept_count [num_lists].
__global__ void __launch_bounds__(256)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const int32_t* __restrict__ counts, int n, int nb,
                int max_keep, int32_t* __restrict__ kept_pos, int32_t* __restrict__ kept_count) {
    __shared__ unsigned long long remv[kNmsMaxN / 64];
    __shared__ unsigned long long diag[64];
    __shared__ unsigned long long s_keep;
    const int l = blockIdx.x;
    const int cnt = counts[l];
    const unsigned long long* mk = mask + (size_t)l * n * nb;
    const int nbv = (cnt + 63) / 64;
    for (int w = threadIdx.x; w < nbv; w += blockDim.x) remv[w] = 0ull;
    __syncthreads();
    int kept = 0;
    for (int b = 0; b < nbv && kept < max_keep; ++b) {
        const int base = b * 64;
        const int bn = min(64, cnt - base);
        if ((int)threadIdx.x < bn) diag[threadIdx.x] = mk[(size_t)(base + threadIdx.x) * nb + b];
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long cur = remv[b], kb = 0ull;
            for (int t = 0; t < bn; ++t) {
                if (!((cur >> t) & 1ull)) {
                    kb |= 1ull << t;
                    cur |= diag[t];
                }
            }
            s_keep = kb;
        }
        __syncthreads();
        const unsigned long long kb = s_keep;
        for (int w = b + 1 + threadIdx.x; w < nbv; w += blockDim.x) {
            unsigned long long acc = remv[w];
            unsigned long long rest = kb;
            while (rest) {
                const int t = __ffsll((long long)rest) - 1;
                rest &= rest - 1;
                acc |= mk[(size_t)(base + t) * nb + w];
            }
            remv[w] = acc;
        }
        if (threadIdx.x < 64 && ((kb >> threadIdx.x) & 1ull)) {
            const int pos = kept + __popcll(kb & ((1ull << threadIdx.x) - 1ull));
            if (pos < max_keep) kept_pos[(size_t)l * max_keep + pos] = base + threadIdx.x;
        }
        kept += __popcll(kb);
        __syncthreads();
    }
    if (threadIdx.x == 0) kept_count[l] = min(kept, max_keep);
}

__global__ void nms_emit_kernel(const int32_t* __restrict__ sorted_idx, const int32_t* __restrict__ kept_pos,
                                const int32_t* __restrict__ kept_count, int64_t* __restrict__ keep,
                                int32_t* __restrict__ num_keep) {
    const int cnt = kept_count[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x)
        keep[i] = (int64_t)sorted_idx[kept_pos[i]];
    if (blockIdx.x == 0 && threadIdx.x == 0) *num_keep = cnt;
}

// rows with any non-finite box coordinate or score are dropped (fast_rcnn_oicr.py:110-114). One warp per row.
__global__ void detect_rowvalid_kernel(const float* __restrict__ probs, const float* __restrict__ pred_boxes, int R,
                                       int C, uint8_t* __restrict__ rowvalid) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    bool ok = true;
    for (int j = lane; j < C + 1; j += 32) ok = ok && isfinite(probs[(size_t)r * (C + 1) + j]);
    for (int j = lane; j < 4 * C; j += 32) ok = ok && isfinite(pred_boxes[(size_t)r * 4 * C + j]);
    ok = __all_sync(FULL_MASK, ok);
    if (lane == 0) rowvalid[r] = ok ? 1 : 0;
}

// Single CTA: merge the per-class score-descending kept lists (first `topk` of each) into the global top-k,
// ordered by (score desc, row-major (r, c) position asc) == batched_nms' final stable sort.
__global__ void __launch_bounds__(kSortThreads)
detect_topk_kernel(const float* __restrict__ probs, const float* __restrict__ pred_boxes, const int32_t* __restrict__ sorted_idx,
                   const int32_t* __restrict__ kept_pos, const int32_t* __restrict__ kept_count, int R, int C, int topk,
                   int n_pow2, float img_h, float img_w, float* __restrict__ det_boxes, float* __restrict__ det_scores,
                   int32_t* __restrict__ det_classes, int32_t* __restrict__ det_rows, int32_t* __restrict__ num_det) {
    extern __shared__ __align__(16) unsigned long long keys[];
    __shared__ int s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        unsigned long long key = 0ull;
        if (i < C * topk) {
            const int c = i / topk, p = i % topk;
            if (p < kept_count[c]) {
                const int r = sorted_idx[(size_t)c * R + kept_pos[(size_t)c * topk + p]];
                const float sc = probs[(size_t)r * (C + 1) + c];
                key = ((unsigned long long)float_to_ordered(sc) << 32) |
                      (unsigned long long)(0xFFFFFFFFu - (unsigned)(r * C + c));
                ++local;
            }
        }
        keys[i] = key;
    }
    local = warp_sum_int(local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_total, local);
    __syncthreads();
    bitonic_sort_desc_u64(keys, n_pow2);
    const int nd = min(s_total, topk);
    if (threadIdx.x == 0) *num_det = nd;
    for (int i = threadIdx.x; i < nd; i += blockDim.x) {
        const unsigned long long key = keys[i];
        const unsigned pos = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
        const int r = pos / C, c = pos % C;
        const float4 b = clip_box(reinterpret_cast<const float4*>(pred_boxes)[(size_t)r * C + c], img_w, img_h);
        reinterpret_cast<float4*>(det_boxes)[i] = b;
        det_scores[i] = probs[(size_t)r * (C + 1) + c];
        det_classes[i] = c;
        det_rows[i] = r;
    }
}

// predict: one warp per row.  probs = (sum_k softmax(logits_k)) / K ; deltas = (sum_k deltas_k) / K ; apply_deltas.
template <int CJ>
__global__ void __launch_bounds__(256)
predict_kernel(const float* __restrict__ logits, long long ld, int col_ref0, int ref_stride, const float* __restrict__ boxes,
               int R, int C, int K, float wx, float wy, float ww, float wh, float* __restrict__ probs,
               float* __restrict__ pred_boxes) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int C1 = C + 1;
    float acc[CJ];
#pragma unroll
    for (int j = 0; j < CJ; ++j) acc[j] = 0.f;
    for (int k = 0; k < K; ++k) {
        const float* row = logits + (size_t)r * ld + col_ref0 + (size_t)k * ref_stride;
        float x[CJ], e[CJ];
        float m = -FLT_MAX;
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            x[j] = (c < C1) ? row[c] : -FLT_MAX;
            m = fmaxf(m, x[j]);
        }
        m = warp_max(m);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            e[j] = (c < C1) ? expf(x[j] - m) : 0.f;
            s += e[j];
        }
        s = warp_sum(s);
#pragma unroll
        for (int j = 0; j < CJ; ++j) acc[j] += e[j] / s;
    }
#pragma unroll
    for (int j = 0; j < CJ; ++j) {
        const int c = lane + 32 * j;
        if (c < C1) probs[(size_t)r * C1 + c] = acc[j] / (float)K;
    }
    // boxes (detectron2/modeling/box_regression.py:73-110)
    const float4 b = reinterpret_cast<const float4*>(boxes)[r];
    const float bw = b.z - b.x, bh = b.w - b.y;
    const float cx = b.x + 0.5f * bw, cy = b.y + 0.5f * bh;
    const float clampv = 4.135166556742356f;  // log(1000/16)
    for (int c = lane; c < C; c += 32) {
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < K; ++k) {
            const float* dr = logits + (size_t)r * ld + col_ref0 + (size_t)k * ref_stride + C1 + 4 * c;
#pragma unroll
            for (int t = 0; t < 4; ++t) d[t] += dr[t];
        }
        const float dx = (d[0] / (float)K) / wx, dy = (d[1] / (float)K) / wy;
        const float dw = fminf((d[2] / (float)K) / ww, clampv), dh = fminf((d[3] / (float)K) / wh, clampv);
        const float pcx = dx * bw + cx, pcy = dy * bh + cy;
        const float pw = expf(dw) * bw, ph = expf(dh) * bh;
        reinterpret_cast<float4*>(pred_boxes)[(size_t)r * C + c] =
            make_float4(pcx - 0.5f * pw, pcy - 0.5f * ph, pcx + 0.5f * pw, pcy + 0.5f * ph);
    }
}

__global__ void tta_accumulate_kernel(const float* __restrict__ pred_boxes, const float* __restrict__ probs, long long nbox,
                                      long long nprob, float sx, float sy, int flipped, float view_w, int first,
                                      float fdiv, float* __restrict__ acc_boxes, float* __restrict__ acc_probs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nbox) {
        float4 b = reinterpret_cast<const float4*>(pred_boxes)[i];
        if (flipped) {
            const float x1 = view_w - b.z, x2 = view_w - b.x;
            b.x = x1;
            b.z = x2;
        }
        b.x *= sx;
        b.z *= sx;
        b.y *= sy;
        b.w *= sy;
        float4 a = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(acc_boxes)[i];
        a.x += b.x;
        a.y += b.y;
        a.z += b.z;
        a.w += b.w;
        if (fdiv > 0.f) {
            a.x /= fdiv;
            a.y /= fdiv;
            a.z /= fdiv;
            a.w /= fdiv;
        }
        reinterpret_cast<float4*>(acc_boxes)[i] = a;
    }
    if (i < nprob) {
        float a = first ? 0.f : acc_probs[i];
        a += probs[i];
        if (fdiv > 0.f) a /= fdiv;
        acc_probs[i] = a;
    }
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct NmsWorkspace {
    int32_t* sorted_idx;
    float4* sorted_box;
    int32_t* counts;
    unsigned long long* mask;
    int32_t* kept_pos;
    int32_t* kept_count;
    uint8_t* rowvalid;
    size_t total;
};

static NmsWorkspace carve(void* base, int lists, int n, int max_keep) {
    NmsWorkspace w;
    const int nb = (n + 63) / 64;
    size_t off = 0;
    uint8_t* p = reinterpret_cast<uint8_t*>(base);
    w.sorted_box = reinterpret_cast<float4*>(p + off);
    off += align256((size_t)lists * n * 16);
    w.mask = reinterpret_cast<unsigned long long*>(p + off);
    off += align256((size_t)lists * n * nb * 8);
    w.sorted_idx = reinterpret_cast<int32_t*>(p + off);
    off += align256((size_t)lists * n * 4);
    w.kept_pos = reinterpret_cast<int32_t*>(p + off);
    off += align256((size_t)lists * max_keep * 4);
    w.counts = reinterpret_cast<int32_t*>(p + off);
    off += align256((size_t)lists * 4);
    w.kept_count = reinterpret_cast<int32_t*>(p + off);
    off += align256((size_t)lists * 4);
    w.rowvalid = p + off;
    off += align256((size_t)n);
    w.total = off;
    return w;
}

constexpr int kDetectMaxTopk = 1024;

static int run_nms_pipeline(const NmsSource& src, int lists, int n, float thr, int max_keep, const NmsWorkspace& w,
                            cudaStream_t st) {
    const int n_pow2 = next_pow2(n < 2 ? 2 : n);
    const int nb = (n + 63) / 64;
    const size_t smem = (size_t)n_pow2 * 8;
    SOSWSOD_ENSURE_SMEM(nms_sort_kernel, smem);
    nms_sort_kernel<<<lists, kSortThreads, smem, st>>>(src, n, n_pow2, w.sorted_idx, w.sorted_box, w.counts);
    SOSWSOD_CHECK_LAUNCH();
    nms_mask_kernel<<<dim3(nb, nb, lists), 64, 0, st>>>(w.sorted_box, w.counts, n, nb, thr, w.mask);
    SOSWSOD_CHECK_LAUNCH();
    nms_scan_kernel<<<lists, 256, 0, st>>>(w.mask, w.counts, n, nb, max_keep, w.kept_pos, w.kept_count);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

}  // namespace soswsod

using namespace soswsod;

extern "C" int soswsod_predict(const float* logits, long long ld, int col_ref0, int ref_stride, const float* boxes,
                               int R, int C, int K, float wx, float wy, float ww, float wh, float* probs,
                               float* pred_boxes, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(logits && boxes && probs && pred_boxes, "predict: null pointer");
    SOSWSOD_CHECK_ARG(R > 0 && C > 0 && K > 0 && C + 1 <= 128, "predict: bad shape");
    SOSWSOD_CHECK_ARG(col_ref0 >= 0 && col_ref0 + (long long)(K - 1) * ref_stride + 5 * C + 1 <= ld, "predict: columns exceed ld");
    SOSWSOD_CHECK_ARG(((uintptr_t)boxes & 15) == 0 && ((uintptr_t)pred_boxes & 15) == 0, "predict: boxes must be 16B aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (R * 32 + 255) / 256;
    const int cj = (C + 1 + 31) / 32;
#define LAUNCH(CJ) predict_kernel<CJ><<<blocks, 256, 0, st>>>(logits, ld, col_ref0, ref_stride, boxes, R, C, K, wx, wy, ww, wh, probs, pred_boxes)
    switch (cj) { case 1: LAUNCH(1); break; case 2: LAUNCH(2); break; case 3: LAUNCH(3); break; default: LAUNCH(4); break; }
#undef LAUNCH
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

extern "C" int soswsod_tta_accumulate(const float* pred_boxes, const float* probs, int R, int C, float scale_x,
                                      float scale_y, int flipped, float view_w, int first, float finalize_div,
                                      float* acc_boxes, float* acc_probs, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(pred_boxes && probs && acc_boxes && acc_probs, "tta_accumulate: null pointer");
    SOSWSOD_CHECK_ARG(R > 0 && C > 0, "tta_accumulate: bad shape");
    SOSWSOD_CHECK_ARG(((uintptr_t)pred_boxes & 15) == 0 && ((uintptr_t)acc_boxes & 15) == 0, "tta_accumulate: boxes must be 16B aligned");
    const long long nbox = (long long)R * C, nprob = (long long)R * (C + 1);
    const long long total = nbox > nprob ? nbox : nprob;
    tta_accumulate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        pred_boxes, probs, nbox, nprob, scale_x, scale_y, flipped, view_w, first, finalize_div, acc_boxes, acc_probs);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

extern "C" size_t soswsod_nms_workspace_bytes(int n) {
    if (n <= 0) return 256;
    return carve(nullptr, 1, n, n).total;
}

extern "C" int soswsod_nms(const float* boxes, const float* scores, int n, float iou_thr, int64_t* keep,
                           int32_t* num_keep, void* workspace, size_t workspace_bytes, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(keep && num_keep, "nms: null output");
    SOSWSOD_CHECK_ARG(n >= 0, "nms: bad n");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        SOSWSOD_CHECK_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int32_t), st));
        return SOSWSOD_OK;
    }
    SOSWSOD_CHECK_ARG(boxes && scores && workspace, "nms: null pointer");
    SOSWSOD_CHECK_ARG(((uintptr_t)boxes & 15) == 0, "nms: boxes must be 16B aligned");
    if (n > kNmsMaxN) {
        set_error("nms: n=%d > %d unsupported", n, kNmsMaxN);
        return SOSWSOD_ERR_UNSUPPORTED;
    }
    if (workspace_bytes < soswsod_nms_workspace_bytes(n)) {
        set_error("nms: workspace too small");
        return SOSWSOD_ERR_WORKSPACE;
    }
    NmsWorkspace w = carve(workspace, 1, n, n);
    NmsSource src{};
    src.mode = 0;
    src.boxes = boxes;
    src.scores = scores;
    int rc = run_nms_pipeline(src, 1, n, iou_thr, n, w, st);
    if (rc) return rc;
    nms_emit_kernel<<<(n + 255) / 256, 256, 0, st>>>(w.sorted_idx, w.kept_pos, w.kept_count, keep, num_keep);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

extern "C" size_t soswsod_detect_workspace_bytes(int R, int C) {
    if (R <= 0 || C <= 0) return 256;
    return carve(nullptr, C, R, kDetectMaxTopk).total;
}

extern "C" int soswsod_detect(const float* probs, const float* pred_boxes, int R, int C, float img_h, float img_w,
                              float score_thr, float nms_thr, int topk, float* det_boxes, float* det_scores,
                              int32_t* det_classes, int32_t* det_rows, int32_t* num_det, void* workspace,
                              size_t workspace_bytes, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(probs && pred_boxes && det_boxes && det_scores && det_classes && det_rows && num_det && workspace,
                      "detect: null pointer");
    SOSWSOD_CHECK_ARG(R > 0 && C > 0 && topk > 0, "detect: bad shape");
    SOSWSOD_CHECK_ARG(((uintptr_t)pred_boxes & 15) == 0 && ((uintptr_t)det_boxes & 15) == 0, "detect: boxes must be 16B aligned");
    if (R > kNmsMaxN || topk > kDetectMaxTopk || (long long)C * topk > 16384) {
        set_error("detect: R=%d (max %d), topk=%d (max %d), C*topk=%lld (max 16384) unsupported", R, kNmsMaxN, topk,
                  kDetectMaxTopk, (long long)C * topk);
        return SOSWSOD_ERR_UNSUPPORTED;
    }
    if (workspace_bytes < soswsod_detect_workspace_bytes(R, C)) {
        set_error("detect: workspace too small");
        return SOSWSOD_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    NmsWorkspace w = carve(workspace, C, R, kDetectMaxTopk);
    detect_rowvalid_kernel<<<(R * 32 + 255) / 256, 256, 0, st>>>(probs, pred_boxes, R, C, w.rowvalid);
    SOSWSOD_CHECK_LAUNCH();
    NmsSource src{};
    src.mode = 1;
    src.boxes = pred_boxes;
    src.scores = probs;
    src.rowvalid = w.rowvalid;
    src.C = C;
    src.img_h = img_h;
    src.img_w = img_w;
    src.score_thr = score_thr;
    // kept_pos rows are laid out with stride kDetectMaxTopk in the workspace but indexed with `topk` below,
    // so run the scan with max_keep = topk and stride topk (fits: topk <= kDetectMaxTopk).
    int rc = run_nms_pipeline(src, C, R, nms_thr, topk, w, st);
    if (rc) return rc;
    const int n_pow2 = next_pow2(C * topk < 2 ? 2 : C * topk);
    const size_t smem = (size_t)n_pow2 * 8;
    SOSWSOD_ENSURE_SMEM(detect_topk_kernel, smem);
    detect_topk_kernel<<<1, kSortThreads, smem, st>>>(probs, pred_boxes, w.sorted_idx, w.kept_pos, w.kept_count, R, C,
                                                     topk, n_pow2, img_h, img_w, det_boxes, det_scores, det_classes,
                                                     det_rows, num_det);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}
