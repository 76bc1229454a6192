// Fused WSDDN two-stream head (kernel (3)): softmax over classes x softmax
// over proposals, image-level scores, BCE, and the gradient w.r.t. both logit blocks.  One thread-block CLUSTER of
// kWsddnCluster CTAs per view: each CTA takes a contiguous slice of the proposals, the three cross-proposal
// reductions (column max, column sum of exponentials, image scores) are combined through distributed shared memory
// in rank order (deterministic), so the whole head stays one launch and the reduction never touches global memory.
//
// Layout inside the CTA: lane <-> class column (c = lane + 32*j), warp <-> proposal row
// (r = warp + 32*i).  Row reductions (softmax over classes) are warp shuffles; column reductions
// (softmax over proposals, image scores) are per-thread partials combined across the 32 warps through
// shared memory -- the cross-proposal reduction never leaves the CTA.
//
// Math restated from uwsod/projects/WSL/wsl/modeling/roi_heads/fast_rcnn_wsddn.py:566-567 (scores),
// :360-375 (image scores, clamp [1e-6, 1-1e-6]) and :340-358 (BCE, mean over C, / N_img = 1).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace soswsod {

constexpr int kWsddnThreads = 1024;
constexpr int kWsddnMaxCJ = 4;  // classes <= 128
constexpr int kWsddnCluster = 8;

template <int CJ>
__global__ void __cluster_dims__(kWsddnCluster, 1, 1) __launch_bounds__(kWsddnThreads, 1)
wsddn_kernel(const float* __restrict__ logits, long long ld, int col_cls, int col_det, int R, int C,
             const float* __restrict__ gt_onehot, float* __restrict__ scores, float* __restrict__ img_scores,
             float* __restrict__ loss, float* __restrict__ dlogits, long long ld_d) {
    __shared__ float red[32][32 * CJ + 1];
    __shared__ float col_max[32 * CJ], col_sum[32 * CJ], col_raw[32 * CJ], col_dp[32 * CJ];
    __shared__ float part[3][32 * CJ];   // this CTA's partial of each of the three reductions (read by its peers)
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int v = blockIdx.x / kWsddnCluster;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* lg = logits + (size_t)v * R * ld;
    float* sc = scores + (size_t)v * R * C;
    const int rows_per = (R + kWsddnCluster - 1) / kWsddnCluster;
    const int r_lo = min(rank * rows_per, R), r_hi = min(r_lo + rows_per, R);
    // value of column c combined over the cluster, in rank order (identical in every CTA)
    auto cluster_gather = [&](int which, int c, bool is_max) {
        float acc = is_max ? -FLT_MAX : 0.f;
        for (int k = 0; k < kWsddnCluster; ++k) {
            const float* peer = cluster.map_shared_rank(&part[which][0], k);
            acc = is_max ? fmaxf(acc, peer[c]) : acc + peer[c];
        }
        return acc;
    };

    // ---- pass 1: column max of the detection logits over proposals ----
    float pm[CJ];
#pragma unroll
    for (int j = 0; j < CJ; ++j) pm[j] = -FLT_MAX;
    for (int r = r_lo + warp; r < r_hi; r += 32) {
        const float* row = lg + (size_t)r * ld + col_det;
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            if (c < C) pm[j] = fmaxf(pm[j], row[c]);
        }
    }
#pragma unroll
    for (int j = 0; j < CJ; ++j) red[warp][lane + 32 * j] = pm[j];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            float m = -FLT_MAX;
            for (int w = 0; w < 32; ++w) m = fmaxf(m, red[w][lane + 32 * j]);
            part[0][lane + 32 * j] = m;
        }
    }
    cluster.sync();
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < CJ; ++j) col_max[lane + 32 * j] = cluster_gather(0, lane + 32 * j, true);
    }
    __syncthreads();

    // ---- pass 2: column sum of exp(det - max) ----
    float ps[CJ];
#pragma unroll
    for (int j = 0; j < CJ; ++j) ps[j] = 0.f;
    for (int r = r_lo + warp; r < r_hi; r += 32) {
        const float* row = lg + (size_t)r * ld + col_det;
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            if (c < C) ps[j] += expf(row[c] - col_max[c]);
        }
    }
#pragma unroll
    for (int j = 0; j < CJ; ++j) red[warp][lane + 32 * j] = ps[j];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            float s = 0.f;
            for (int w = 0; w < 32; ++w) s += red[w][lane + 32 * j];
            part[1][lane + 32 * j] = s;
        }
    }
    cluster.sync();
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < CJ; ++j) col_sum[lane + 32 * j] = cluster_gather(1, lane + 32 * j, false);
    }
    __syncthreads();

    // ---- pass 3: scores = softmax_c(cls) * softmax_r(det); column sums of the scores ----
#pragma unroll
    for (int j = 0; j < CJ; ++j) ps[j] = 0.f;
    for (int r = r_lo + warp; r < r_hi; r += 32) {
        const float* rc = lg + (size_t)r * ld + col_cls;
        const float* rd = lg + (size_t)r * ld + col_det;
        float x[CJ], e[CJ];
        float m = -FLT_MAX;
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            x[j] = (c < C) ? rc[c] : -FLT_MAX;
            m = fmaxf(m, x[j]);
        }
        m = warp_max(m);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            e[j] = (c < C) ? expf(x[j] - m) : 0.f;
            s += e[j];
        }
        s = warp_sum(s);
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            if (c < C) {
                const float a = e[j] / s;
                const float bb = expf(rd[c] - col_max[c]) / col_sum[c];
                const float val = a * bb;
                sc[(size_t)r * C + c] = val;
                ps[j] += val;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CJ; ++j) red[warp][lane + 32 * j] = ps[j];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            float s = 0.f;
            for (int w = 0; w < 32; ++w) s += red[w][c];
            part[2][c] = s;
        }
    }
    cluster.sync();
    if (warp == 0) {
        float lsum = 0.f;
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            const float s = cluster_gather(2, c, false);
            float dp = 0.f;
            if (c < C) {
                const float p = fminf(fmaxf(s, 1e-6f), 1.0f - 1e-6f);
                const float t = gt_onehot[c];
                if (rank == 0) img_scores[(size_t)v * C + c] = p;
                // torch BCE clamps each log term at -100
                const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
                lsum += -(t * lp + (1.f - t) * l1p);
                const bool inside = (s >= 1e-6f) && (s <= 1.0f - 1e-6f);  // clamp passes gradient inside [min,max]
                dp = inside ? ((-t / p + (1.f - t) / (1.f - p)) / (float)C) : 0.f;
            }
            col_raw[c] = s;
            col_dp[c] = dp;
        }
        lsum = warp_sum(lsum);
        if (lane == 0 && rank == 0) loss[v] = lsum / (float)C;
    }
    cluster.sync();   // peers have read part[] before any CTA of the cluster may exit
    if (dlogits == nullptr) return;

    // ---- pass 4: gradients.  S = A*B, A = softmax_c(cls), B = softmax_r(det), dS_rc = dp_c:
    //   dCls_rc = A_rc * (dp_c*B_rc - sum_c' dp_c'*S_rc')     dDet_rc = dp_c * B_rc * (A_rc - rawsum_c)
    float* dg = dlogits + (size_t)v * R * ld_d;
    for (int r = r_lo + warp; r < r_hi; r += 32) {
        const float* rc = lg + (size_t)r * ld + col_cls;
        const float* rd = lg + (size_t)r * ld + col_det;
        float x[CJ], e[CJ];
        float m = -FLT_MAX;
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            x[j] = (c < C) ? rc[c] : -FLT_MAX;
            m = fmaxf(m, x[j]);
        }
        m = warp_max(m);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            e[j] = (c < C) ? expf(x[j] - m) : 0.f;
            s += e[j];
        }
        s = warp_sum(s);
        float a[CJ], bb[CJ];
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            a[j] = 0.f;
            bb[j] = 0.f;
            if (c < C) {
                a[j] = e[j] / s;
                bb[j] = expf(rd[c] - col_max[c]) / col_sum[c];
                dot += col_dp[c] * a[j] * bb[j];
            }
        }
        dot = warp_sum(dot);
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            if (c < C) {
                dg[(size_t)r * ld_d + col_cls + c] = a[j] * (col_dp[c] * bb[j] - dot);
                dg[(size_t)r * ld_d + col_det + c] = col_dp[c] * bb[j] * (a[j] - col_raw[c]);
            }
        }
    }
}

}  // namespace soswsod

using namespace soswsod;

extern "C" int soswsod_wsddn_forward(const float* logits, long long ld, int col_cls, int col_det, int num_views,
                                     int R, int C, const float* gt_onehot, float* scores, float* img_scores,
                                     float* loss, float* dlogits, long long ld_d, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(logits && gt_onehot && scores && img_scores && loss, "wsddn_forward: null pointer");
    SOSWSOD_CHECK_ARG(num_views > 0 && R > 0 && C > 0, "wsddn_forward: bad shape");
    SOSWSOD_CHECK_ARG(C <= 32 * kWsddnMaxCJ, "wsddn_forward: C=%d > %d unsupported", C, 32 * kWsddnMaxCJ);
    SOSWSOD_CHECK_ARG(col_cls >= 0 && col_det >= 0 && col_cls + C <= ld && col_det + C <= ld, "wsddn_forward: bad columns");
    SOSWSOD_CHECK_ARG(!dlogits || (col_cls + C <= ld_d && col_det + C <= ld_d), "wsddn_forward: bad ld_d");
    cudaStream_t st = (cudaStream_t)stream;
    const int cj = (C + 31) / 32;
    switch (cj) {
        case 1: wsddn_kernel<1><<<num_views * kWsddnCluster, kWsddnThreads, 0, st>>>(logits, ld, col_cls, col_det, R, C, gt_onehot, scores, img_scores, loss, dlogits, ld_d); break;
        case 2: wsddn_kernel<2><<<num_views * kWsddnCluster, kWsddnThreads, 0, st>>>(logits, ld, col_cls, col_det, R, C, gt_onehot, scores, img_scores, loss, dlogits, ld_d); break;
        case 3: wsddn_kernel<3><<<num_views * kWsddnCluster, kWsddnThreads, 0, st>>>(logits, ld, col_cls, col_det, R, C, gt_onehot, scores, img_scores, loss, dlogits, ld_d); break;
        default: wsddn_kernel<4><<<num_views * kWsddnCluster, kWsddnThreads, 0, st>>>(logits, ld, col_cls, col_det, R, C, gt_onehot, scores, img_scores, loss, dlogits, ld_d); break;
    }
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}
