// OICR refinement branches (kernel (4), SURVEY.md §8a rows G..P) for all K branches at once:
//   avg_scores   -- view-averaged detached scores each branch mines from (row G)
//   topk         -- per-GT-class top-k seed candidates, bitonic sort in shared memory (row H)
//   nms_label    -- class-agnostic greedy NMS of the candidates + IoU/argmax/threshold labelling of every
//                   proposal, all in one CTA per branch with the IoU evaluated on the fly (rows I..L)
//   loss         -- weighted CE + L1 box regression + gradients + accuracy counters (rows N, O, P)
//
// Integer results (seed indices/classes, labels, matched indices) are bit-exact w.r.t. the reference
// semantics given identical fp32 scores: IoU uses explicit round-to-nearest mul/add/sub/div (no FMA),
// thresholds are compared in fp32, max over seeds keeps the first maximum, ordering is
// (score desc, index asc).  Reference lines are cited per kernel.
#include <cooperative_groups.h>

#include "common.cuh"

namespace soswsod {

constexpr int kOicrThreads = 1024;
constexpr int kMaxCJ = 4;  // (C+1) <= 128

// -------------------------------------------------------------------------------------------------
// avg_scores: roi_heads_oicrplus.py:290-294 (k = 0) and :390-395 + fast_rcnn_oicr.py:702-716 (k >= 1)
// grid (ceil(R/32), K), block 1024 = 32 warps, one warp per row.
// -------------------------------------------------------------------------------------------------
template <int CJ>
__global__ void __launch_bounds__(kOicrThreads)
oicr_avg_scores_kernel(const float* __restrict__ wsddn_scores, const float* __restrict__ logits, long long ld,
                       int col_ref0, int ref_stride, int V, int R, int C, float* __restrict__ prev) {
    const int k = blockIdx.y;
    const int r = blockIdx.x * 32 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int C1 = C + 1;
    float* out = prev + ((size_t)k * R + r) * C1;
    float acc[CJ];
#pragma unroll
    for (int j = 0; j < CJ; ++j) acc[j] = 0.f;
    if (k == 0) {
        for (int v = 0; v < V; ++v) {
            const float* row = wsddn_scores + ((size_t)v * R + r) * C;
#pragma unroll
            for (int j = 0; j < CJ; ++j) {
                const int c = lane + 32 * j;
                if (c < C) acc[j] = (v == 0) ? row[c] : __fadd_rn(acc[j], row[c]);
            }
        }
    } else {
        const int col = col_ref0 + (k - 1) * ref_stride;
        for (int v = 0; v < V; ++v) {
            const float* row = logits + ((size_t)v * R + r) * ld + col;
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
            for (int j = 0; j < CJ; ++j) {
                const float p = e[j] / s;
                acc[j] = (v == 0) ? p : __fadd_rn(acc[j], p);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CJ; ++j) {
        const int c = lane + 32 * j;
        if (c < C1) out[c] = (k == 0 && c == C) ? 0.f : __fdiv_rn(acc[j], (float)V);
    }
}

// -------------------------------------------------------------------------------------------------
// topk: roi_heads_oicrplus.py:646-669 (index_select of the GT columns, topk over proposals) and
// :698-704 (score >= thres with rank 0 forced).  grid (G, K); dynamic smem = next_pow2(R) u64 keys.
// Candidate slot = rank*G + g  (the reference's row-major masked_select order, :705-731).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kOicrThreads)
oicr_topk_kernel(const float* __restrict__ prev, long long ld_prev, const int32_t* __restrict__ gt_classes, int G,
                 const int32_t* __restrict__ gt_count, int R, int kt, float score_thr, int n_pow2,
                 unsigned long long* __restrict__ cand_key, int32_t* __restrict__ cand_row) {
    extern __shared__ __align__(16) unsigned long long keys[];
    const int g = blockIdx.x, k = blockIdx.y;
    // G is the capacity of gt_classes; with a device-side count only the first *gt_count entries are live (the
    // slots of the others stay empty -- candidate order is rank-major then g either way)
    if (gt_count && g >= *gt_count) {
        unsigned long long* ck0 = cand_key + (size_t)k * kt * G;
        for (int i = threadIdx.x; i < kt; i += blockDim.x) ck0[(size_t)i * G + g] = 0ull;
        return;
    }
    const int cls = gt_classes[g];
    const float* col = prev + (size_t)k * R * ld_prev + cls;
    for (int r = threadIdx.x; r < n_pow2; r += blockDim.x) {
        unsigned long long key = 0ull;
        if (r < R) key = ((unsigned long long)float_to_ordered(col[(size_t)r * ld_prev]) << 32) | (0xFFFFFFFFu - (unsigned)r);
        keys[r] = key;
    }
    __syncthreads();
    bitonic_sort_desc_u64(keys, n_pow2);
    unsigned long long* ck = cand_key + (size_t)k * kt * G;
    int32_t* cr = cand_row + (size_t)k * kt * G;
    for (int i = threadIdx.x; i < kt; i += blockDim.x) {
        const unsigned long long key = keys[i];
        const float score = ordered_to_float((uint32_t)(key >> 32));
        const unsigned row = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
        const bool in = (i == 0) || (score >= score_thr);
        const unsigned pos = (unsigned)i * G + g;
        ck[pos] = in ? ((key & 0xFFFFFFFF00000000ull) | (0xFFFFFFFFu - pos)) : 0ull;
        cr[pos] = (int32_t)row;
    }
}

__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(FULL_MASK, v, o);
        v = t > v ? t : v;
    }
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    unsigned long long r = (lane < nw) ? scratch[lane] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(FULL_MASK, r, o);
        r = t > r ? t : r;
    }
    __syncthreads();
    return r;
}

// -------------------------------------------------------------------------------------------------
// nms_label: get_pgt_mist NMS step (roi_heads_oicrplus.py:575-603: batched_nms with all-zero idxs ==
// class-agnostic torchvision nms, IoU > thr strict), then pairwise_iou (detectron2/structures/boxes.py:
// 329-361), Matcher (detectron2/modeling/matcher.py:100-106) and the label/weight/index gather of
// wsl/modeling/roi_heads/roi_heads.py:248-252, 327-357.  grid K.
// Greedy NMS without a second sort: repeatedly take the best surviving key, suppress what it overlaps.
// -------------------------------------------------------------------------------------------------
constexpr int kSeedCache = 1024;

__global__ void __launch_bounds__(kOicrThreads)
oicr_nms_label_kernel(unsigned long long* __restrict__ cand_key, const int32_t* __restrict__ cand_row,
                      const float* __restrict__ boxes, const int32_t* __restrict__ gt_classes, int G, int R, int C,
                      int kt, float nms_thr, float iou_lo, float iou_hi, int32_t* __restrict__ seed_count,
                      int32_t* __restrict__ seed_index, int32_t* __restrict__ seed_class,
                      float* __restrict__ seed_score, int32_t* __restrict__ gt_class, float* __restrict__ gt_weight,
                      int32_t* __restrict__ gt_index, int32_t* __restrict__ counts) {
    __shared__ unsigned long long scratch[32];
    __shared__ float4 s_box[kSeedCache];
    __shared__ float s_area[kSeedCache];
    __shared__ int s_cnt[3];
    const int k = blockIdx.x;
    const int M0 = kt * G;
    unsigned long long* ck = cand_key + (size_t)k * M0;
    const int32_t* cr = cand_row + (size_t)k * M0;
    int32_t* sidx = seed_index + (size_t)k * M0;
    int32_t* scls = seed_class + (size_t)k * M0;
    float* ssc = seed_score + (size_t)k * M0;
    const float4* b4 = reinterpret_cast<const float4*>(boxes);

    unsigned long long local = 0ull;
    for (int p = threadIdx.x; p < M0; p += blockDim.x) {
        const unsigned long long key = ck[p];
        local = key > local ? key : local;
    }
    unsigned long long best = block_max_u64(local, scratch);
    int M = 0;
    while (best != 0ull) {
        const unsigned pos = 0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull);
        const int row = cr[pos];
        const float4 kb = b4[row];
        if (threadIdx.x == 0) {
            sidx[M] = row;
            scls[M] = gt_classes[pos % G];
            ssc[M] = ordered_to_float((uint32_t)(best >> 32));
            if (M < kSeedCache) {
                s_box[M] = kb;
                s_area[M] = box_area_rn(kb.x, kb.y, kb.z, kb.w);
            }
        }
        ++M;
        local = 0ull;
        for (int p = threadIdx.x; p < M0; p += blockDim.x) {
            const unsigned long long key = ck[p];
            if (key == 0ull) continue;
            if ((unsigned)p == pos) {
                ck[p] = 0ull;
                continue;
            }
            const float iou = box_iou_nms_rn(kb, b4[cr[p]]);
            if (iou > nms_thr)
                ck[p] = 0ull;
            else
                local = key > local ? key : local;
        }
        best = block_max_u64(local, scratch);  // contains the barriers that order the ck[] updates
    }
    if (threadIdx.x == 0) seed_count[k] = M;
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();  // seeds (global + smem) visible to the whole CTA

    int n_fg = 0, n_bg = 0, n_ig = 0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const float4 pb = b4[r];
        const float pa = box_area_rn(pb.x, pb.y, pb.z, pb.w);
        float bv = -1.f;
        int bm = 0;
        for (int m = 0; m < M; ++m) {
            float4 sb;
            float sa;
            if (m < kSeedCache) {
                sb = s_box[m];
                sa = s_area[m];
            } else {
                sb = b4[sidx[m]];
                sa = box_area_rn(sb.x, sb.y, sb.z, sb.w);
            }
            const float v = box_iou_pairwise_rn(sb, sa, pb, pa);
            if (v > bv) {
                bv = v;
                bm = m;
            }
        }
        // Matcher: labels [0,-1,1] over [-inf,lo), [lo,hi), [hi,inf); initial label 1
        int lab = 1;
        if (bv < iou_lo) lab = 0;
        if (bv >= iou_lo && bv < iou_hi) lab = -1;
        int y = scls[bm];
        if (lab == 0) y = C;
        if (lab == -1) y = -1;
        gt_class[(size_t)k * R + r] = y;
        gt_weight[(size_t)k * R + r] = ssc[bm];
        gt_index[(size_t)k * R + r] = sidx[bm];
        n_fg += (lab == 1);
        n_bg += (lab == 0);
        n_ig += (lab == -1);
    }
    n_fg = warp_sum_int(n_fg);
    n_bg = warp_sum_int(n_bg);
    n_ig = warp_sum_int(n_ig);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_cnt[0], n_fg);
        atomicAdd(&s_cnt[1], n_bg);
        atomicAdd(&s_cnt[2], n_ig);
    }
    __syncthreads();
    if (threadIdx.x < 3) counts[k * 3 + threadIdx.x] = s_cnt[threadIdx.x];
}

// -------------------------------------------------------------------------------------------------
// loss: OICROutputs (fast_rcnn_oicr.py:157-352) for branch k, logit view u.  grid (V, K), one warp/row.
//   CE: w[y==-1]=0 (:219-220); mean over ALL R rows of CE(ignore_index=-1)*w (:258-273)
//   box: fg rows only, class-specific 4 columns, L1 (smooth_l1 beta=0), / R (:276-352), targets from
//        Box2BoxTransform.get_deltas (detectron2/modeling/box_regression.py:38-71)
//   the branch loss is the mean over the V views (roi_heads_oicrplus.py:384-388); with flip_quirk the last
//   view's loss is evaluated on view V-2's predictions (:381).
// -------------------------------------------------------------------------------------------------
constexpr int kLossCluster = 8;   // CTAs (row slices) per (logit view, branch); partial sums meet in rank order over DSMEM

template <int CJ>
__global__ void __cluster_dims__(kLossCluster, 1, 1) __launch_bounds__(kOicrThreads)
oicr_loss_kernel(const float* __restrict__ logits, long long ld, int col_ref0, int ref_stride,
                 const float* __restrict__ boxes, const int32_t* __restrict__ gt_class,
                 const float* __restrict__ gt_weight, const int32_t* __restrict__ gt_index, int V, int R, int C,
                 int flip_quirk, float wx, float wy, float ww, float wh, float* __restrict__ view_losses,
                 int32_t* __restrict__ acc_counts, float* __restrict__ dlogits, long long ld_d) {
    __shared__ float s_loss[32][4];
    __shared__ int s_acc[32][4];
    __shared__ float s_part[4];   // this CTA's (ce, box0, box1) -- read by rank 0 of the cluster
    __shared__ int s_parti[4];
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int u = blockIdx.x / kLossCluster, k = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows_per = (R + kLossCluster - 1) / kLossCluster;
    const int r_lo = min(rank * rows_per, R), r_hi = min(r_lo + rows_per, R);
    const int C1 = C + 1;
    const int col = col_ref0 + k * ref_stride;
    const bool quirk = flip_quirk && V >= 2;
    // loss views evaluated on this logit view
    int lv[2];
    int nlv = 0;
    if (quirk) {
        if (u == V - 2) { lv[0] = V - 2; lv[1] = V - 1; nlv = 2; }
        else if (u == V - 1) { nlv = 0; }
        else { lv[0] = u; nlv = 1; }
    } else {
        lv[0] = u; nlv = 1;
    }
    const float invR = 1.f / (float)R, invV = 1.f / (float)V;
    const int32_t* yk = gt_class + (size_t)k * R;
    const float* wk = gt_weight + (size_t)k * R;
    const int32_t* gk = gt_index + (size_t)k * R;

    float ce_sum = 0.f;            // per-view CE sum (identical for every loss view of this logit view)
    float box_sum[2] = {0.f, 0.f};
    int n_fg = 0, n_acc = 0, n_fgacc = 0, n_fn = 0;

    for (int r = r_lo + warp; r < r_hi; r += 32) {
        const float* zrow = logits + ((size_t)u * R + r) * ld + col;
        const int y = yk[r];
        const float w = (y == -1) ? 0.f : wk[r];
        float x[CJ], e[CJ];
        float m = -FLT_MAX;
        int am = 0;
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            x[j] = (c < C1) ? zrow[c] : -FLT_MAX;
            if (x[j] > m) { m = x[j]; am = c; }
        }
        // warp arg-max (first maximum wins)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(FULL_MASK, m, o);
            const int oa = __shfl_xor_sync(FULL_MASK, am, o);
            if (om > m || (om == m && oa < am)) { m = om; am = oa; }
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            e[j] = (c < C1) ? expf(x[j] - m) : 0.f;
            s += e[j];
        }
        s = warp_sum(s);
        const float lse = logf(s) + m;
        float zy = 0.f;
        if (y >= 0) {
            const int jy = y >> 5;
            float cand = 0.f;
#pragma unroll
            for (int j = 0; j < CJ; ++j) if (j == jy) cand = x[j];
            zy = __shfl_sync(FULL_MASK, cand, y & 31);
        }
        const bool fg = (y >= 0) && (y < C);
        if (lane == 0 && nlv > 0) {
            if (y >= 0) ce_sum += (lse - zy) * w;
            n_fg += fg;
            n_acc += (am == y);
            n_fgacc += (fg && am == y);
            n_fn += (fg && am == C);
        }
        if (dlogits) {
            float* drow = dlogits + ((size_t)u * R + r) * ld_d + col;
            const float gscale = (y >= 0) ? (w * invR * invV * (float)nlv) : 0.f;
#pragma unroll
            for (int j = 0; j < CJ; ++j) {
                const int c = lane + 32 * j;
                if (c < C1) drow[c] = gscale * (e[j] / s - ((c == y) ? 1.f : 0.f));
            }
            for (int c = lane; c < 4 * C; c += 32) drow[C1 + c] = 0.f;
        }
        if (fg && nlv > 0) {
            __syncwarp();
            const float* drow_in = zrow + C1 + 4 * y;
            float gacc = 0.f;
            for (int t = 0; t < nlv; ++t) {
                const float* bv = boxes + (size_t)lv[t] * R * 4;
                const float sx1 = bv[r * 4 + 0], sy1 = bv[r * 4 + 1], sx2 = bv[r * 4 + 2], sy2 = bv[r * 4 + 3];
                const int gi = gk[r];
                const float tx1 = bv[gi * 4 + 0], ty1 = bv[gi * 4 + 1], tx2 = bv[gi * 4 + 2], ty2 = bv[gi * 4 + 3];
                const float sw = sx2 - sx1, sh = sy2 - sy1;
                const float scx = sx1 + 0.5f * sw, scy = sy1 + 0.5f * sh;
                const float tw = tx2 - tx1, th = ty2 - ty1;
                const float tcx = tx1 + 0.5f * tw, tcy = ty1 + 0.5f * th;
                float tgt = 0.f;
                if (lane == 0) tgt = wx * (tcx - scx) / sw;
                else if (lane == 1) tgt = wy * (tcy - scy) / sh;
                else if (lane == 2) tgt = ww * logf(tw / sw);
                else if (lane == 3) tgt = wh * logf(th / sh);
                float diff = 0.f;
                if (lane < 4) diff = drow_in[lane] - tgt;
                float l1 = fabsf(diff);
                l1 += __shfl_xor_sync(FULL_MASK, l1, 1);
                l1 += __shfl_xor_sync(FULL_MASK, l1, 2);
                if (lane == 0) box_sum[t] += l1;
                gacc += (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
            }
            if (dlogits && lane < 4)
                dlogits[((size_t)u * R + r) * ld_d + col + C1 + 4 * y + lane] = gacc * invR * invV;
        }
    }
    // block reduction over the 32 warps (lane 0 of each warp holds the partials), fixed order
    if (lane == 0) {
        s_loss[warp][0] = ce_sum; s_loss[warp][1] = box_sum[0]; s_loss[warp][2] = box_sum[1]; s_loss[warp][3] = 0.f;
        s_acc[warp][0] = n_fg; s_acc[warp][1] = n_acc; s_acc[warp][2] = n_fgacc; s_acc[warp][3] = n_fn;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ce = 0.f, b0 = 0.f, b1 = 0.f;
        int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int wi = 0; wi < 32; ++wi) {
            ce += s_loss[wi][0]; b0 += s_loss[wi][1]; b1 += s_loss[wi][2];
            a0 += s_acc[wi][0]; a1 += s_acc[wi][1]; a2 += s_acc[wi][2]; a3 += s_acc[wi][3];
        }
        s_part[0] = ce; s_part[1] = b0; s_part[2] = b1;
        s_parti[0] = a0; s_parti[1] = a1; s_parti[2] = a2; s_parti[3] = a3;
    }
    cluster.sync();
    if (threadIdx.x == 0 && nlv > 0 && rank == 0) {
        float ce = 0.f, b0 = 0.f, b1 = 0.f;
        int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int q = 0; q < kLossCluster; ++q) {
            const float* pf = cluster.map_shared_rank(&s_part[0], q);
            const int* pi = cluster.map_shared_rank(&s_parti[0], q);
            ce += pf[0]; b0 += pf[1]; b1 += pf[2];
            a0 += pi[0]; a1 += pi[1]; a2 += pi[2]; a3 += pi[3];
        }
        for (int t = 0; t < nlv; ++t) {
            const int v = lv[t];
            view_losses[((size_t)k * V + v) * 2 + 0] = ce * invR;
            view_losses[((size_t)k * V + v) * 2 + 1] = (t == 0 ? b0 : b1) * invR;
            if (acc_counts) {
                int32_t* a = acc_counts + ((size_t)k * V + v) * 5;
                a[0] = R; a[1] = a0; a[2] = a1; a[3] = a2; a[4] = a3;
            }
        }
    }
    cluster.sync();   // rank 0 has read every peer's partials before any CTA of the cluster exits
}

__global__ void oicr_finalize_kernel(const float* __restrict__ view_losses, int V, int K, float* __restrict__ losses) {
    const int i = threadIdx.x;  // k*2 + which
    if (i >= K * 2) return;
    const int k = i >> 1, which = i & 1;
    float s = 0.f;
    for (int v = 0; v < V; ++v) s += view_losses[((size_t)k * V + v) * 2 + which];
    losses[i] = s / (float)V;
}

}  // namespace soswsod

using namespace soswsod;

extern "C" int soswsod_oicr_avg_scores(const float* wsddn_scores, const float* logits, long long ld, int col_ref0,
                                       int ref_stride, int num_views, int R, int C, int K, float* prev,
                                       soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(wsddn_scores && prev && (K <= 1 || logits), "oicr_avg_scores: null pointer");
    SOSWSOD_CHECK_ARG(num_views > 0 && R > 0 && C > 0 && K > 0, "oicr_avg_scores: bad shape");
    SOSWSOD_CHECK_ARG(C + 1 <= 32 * kMaxCJ, "oicr_avg_scores: C=%d unsupported", C);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((R + 31) / 32, K);
    const int cj = (C + 1 + 31) / 32;
#define LAUNCH(CJ) oicr_avg_scores_kernel<CJ><<<grid, kOicrThreads, 0, st>>>(wsddn_scores, logits, ld, col_ref0, ref_stride, num_views, R, C, prev)
    switch (cj) { case 1: LAUNCH(1); break; case 2: LAUNCH(2); break; case 3: LAUNCH(3); break; default: LAUNCH(4); break; }
#undef LAUNCH
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

namespace soswsod {
// get_image_level_gt (wsl/modeling/roi_heads/roi_heads.py:144-164) without the host round trip of torch.unique:
// sorted distinct classes (padded to C with -1), their count and the one-hot row, all left on the device.
template <typename T>
__global__ void __launch_bounds__(128)
image_level_gt_kernel(const T* __restrict__ gt, int n, int C, int32_t* __restrict__ list, int32_t* __restrict__ count,
                      float* __restrict__ onehot) {
    __shared__ int present[128];
    __shared__ int wtot[4];
    const int c = threadIdx.x;
    present[c] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 128) {
        const long long v = (long long)gt[i];
        if (v >= 0 && v < C) present[(int)v] = 1;   // benign race: every writer stores 1
    }
    __syncthreads();
    const bool on = c < C && present[c];
    const unsigned bal = __ballot_sync(FULL_MASK, on);
    if ((c & 31) == 0) wtot[c >> 5] = __popc(bal);
    __syncthreads();
    int before = __popc(bal & ((1u << (c & 31)) - 1u));
    for (int w = 0; w < (c >> 5); ++w) before += wtot[w];
    const int total = wtot[0] + wtot[1] + wtot[2] + wtot[3];
    if (c < C) {
        onehot[c] = on ? 1.f : 0.f;
        if (on) list[before] = c;
        if (c >= total) list[c] = -1;
    }
    if (c == 0) *count = total;
}
}  // namespace soswsod

extern "C" int soswsod_image_level_gt(const void* gt_classes, int is_int64, int n, int C, int32_t* gt_list,
                                      int32_t* gt_count, float* gt_onehot, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(gt_list && gt_count && gt_onehot && (gt_classes || n == 0), "image_level_gt: null pointer");
    SOSWSOD_CHECK_ARG(n >= 0 && C > 0 && C <= 128, "image_level_gt: need 0 < C <= 128");
    if (is_int64)
        image_level_gt_kernel<long long><<<1, 128, 0, (cudaStream_t)stream>>>((const long long*)gt_classes, n, C, gt_list,
                                                                             gt_count, gt_onehot);
    else
        image_level_gt_kernel<int32_t><<<1, 128, 0, (cudaStream_t)stream>>>((const int32_t*)gt_classes, n, C, gt_list,
                                                                           gt_count, gt_onehot);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

extern "C" size_t soswsod_oicr_mine_workspace_bytes(int top_k, int G, int K) {
    const size_t m0 = (size_t)top_k * G;
    return (size_t)K * m0 * (8 + 4) + 256;
}

extern "C" int soswsod_oicr_mine_label(const float* prev, long long ld_prev, const float* boxes,
                                       const int32_t* gt_classes, int G, const int32_t* gt_count, int R, int C, int K,
                                       int top_k,
                                       float score_thr, float nms_thr, float iou_lo, float iou_hi,
                                       int32_t* seed_count, int32_t* seed_index, int32_t* seed_class,
                                       float* seed_score, int32_t* gt_class, float* gt_weight, int32_t* gt_index,
                                       int32_t* counts, void* workspace, size_t workspace_bytes,
                                       soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(prev && boxes && gt_classes && seed_count && seed_index && seed_class && seed_score &&
                          gt_class && gt_weight && gt_index && counts && workspace, "oicr_mine_label: null pointer");
    SOSWSOD_CHECK_ARG(G >= 1, "oicr_mine_label: image has no GT class (G=%d); the reference cannot label it either", G);
    SOSWSOD_CHECK_ARG(R > 0 && C > 0 && K > 0 && G <= C && top_k >= 1 && top_k <= R, "oicr_mine_label: bad shape");
    SOSWSOD_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "oicr_mine_label: boxes must be 16B aligned");
    const int n_pow2 = next_pow2(R);
    if (n_pow2 > 16384) {
        set_error("oicr_mine_label: R=%d > 16384 unsupported", R);
        return SOSWSOD_ERR_UNSUPPORTED;
    }
    if (workspace_bytes < soswsod_oicr_mine_workspace_bytes(top_k, G, K)) {
        set_error("oicr_mine_label: workspace too small");
        return SOSWSOD_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int kt = top_k;
    const size_t m0 = (size_t)kt * G;
    unsigned long long* cand_key = reinterpret_cast<unsigned long long*>(workspace);
    int32_t* cand_row = reinterpret_cast<int32_t*>(cand_key + (size_t)K * m0);
    const size_t smem = (size_t)n_pow2 * 8;
    SOSWSOD_ENSURE_SMEM(oicr_topk_kernel, smem);
    oicr_topk_kernel<<<dim3(G, K), kOicrThreads, smem, st>>>(prev, ld_prev, gt_classes, G, gt_count, R, kt, score_thr,
                                                            n_pow2, cand_key, cand_row);
    SOSWSOD_CHECK_LAUNCH();
    oicr_nms_label_kernel<<<K, kOicrThreads, 0, st>>>(cand_key, cand_row, boxes, gt_classes, G, R, C, kt, nms_thr,
                                                      iou_lo, iou_hi, seed_count, seed_index, seed_class,
                                                      seed_score, gt_class, gt_weight, gt_index, counts);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

extern "C" int soswsod_oicr_loss(const float* logits, long long ld, int col_ref0, int ref_stride,
                                 const float* boxes, const int32_t* gt_class, const float* gt_weight,
                                 const int32_t* gt_index, int num_views, int R, int C, int K, int flip_quirk,
                                 float wx, float wy, float ww, float wh, float* losses, float* view_losses,
                                 int32_t* acc_counts, float* dlogits, long long ld_d, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(logits && boxes && gt_class && gt_weight && gt_index && losses && view_losses,
                      "oicr_loss: null pointer (view_losses is required scratch)");
    SOSWSOD_CHECK_ARG(num_views > 0 && R > 0 && C > 0 && K > 0 && K <= 512, "oicr_loss: bad shape");
    SOSWSOD_CHECK_ARG(C + 1 <= 32 * kMaxCJ, "oicr_loss: C=%d unsupported", C);
    SOSWSOD_CHECK_ARG(col_ref0 >= 0 && col_ref0 + (long long)(K - 1) * ref_stride + 5 * C + 1 <= ld, "oicr_loss: columns exceed ld");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(num_views * kLossCluster, K);
    const int cj = (C + 1 + 31) / 32;
#define LAUNCH(CJ) oicr_loss_kernel<CJ><<<grid, kOicrThreads, 0, st>>>(logits, ld, col_ref0, ref_stride, boxes, gt_class, gt_weight, gt_index, num_views, R, C, flip_quirk, wx, wy, ww, wh, view_losses, acc_counts, dlogits, ld_d)
    switch (cj) { case 1: LAUNCH(1); break; case 2: LAUNCH(2); break; case 3: LAUNCH(3); break; default: LAUNCH(4); break; }
#undef LAUNCH
    SOSWSOD_CHECK_LAUNCH();
    oicr_finalize_kernel<<<1, 1024, 0, st>>>(view_losses, num_views, K, losses);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}
