// bf16 GEMM on the 5th-gen tensor cores (kernel (2) of the hot path):
//     D[M,N] = epilogue( sum_k A[m,k] * B[n,k] ),  fp32 accumulation in TMEM.
//
// Replaces nn.Linear + F.relu_ + F.dropout of DiscriminativeAdaptionNeck.forward
// (uwsod/projects/WSL/wsl/modeling/roi_heads/box_head.py:82-91), the cls/det and cls_score/bbox_pred
// Linear layers (fast_rcnn_wsddn.py:558-559, fast_rcnn_oicr.py:517-519) and their autograd backward
// (dgrad: D = dY * W, wgrad: D = dY^T * X) -- one kernel, three operand-major combinations.
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor loads of the A / B k-slices into a 4-stage
//              128B-swizzled shared-memory ring, completion on "full" mbarriers
//   warp 1   : MMA issuer    -- one lane issues tcgen05.mma (128 x BLOCK_N x 16, cta_group::1); smem slots are
//              released with tcgen05.commit on the "empty" mbarriers; TMEM holds two accumulator stages
//              so the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2-5: epilogue      -- tcgen05.ld (32 lanes x 32 columns per warp), bias / ReLU / backward mask /
//              dropout, 128-bit stores of fp32 or bf16
// Operand layouts in shared memory are the canonical UMMA layouts: K-major SWIZZLE_128B (rows of 64 bf16 =
// 128 B, 8-row swizzle atoms, SBO = 1024 B) or MN-major SWIZZLE_128B (64 MN elements contiguous per k row,
// 8-k-row atoms, SBO = 1024 B, LBO = BLOCK_K*128 B between 64-wide MN chunks).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"

namespace soswsod {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;   // 64 bf16 = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 192;
constexpr int kNumAccStages = 2;

__host__ __device__ constexpr int gemm_stages(int block_n) { return block_n == 256 ? 4 : 6; }
__host__ __device__ constexpr int gemm_stage_bytes(int block_n) { return (kBlockM + block_n) * kBlockK * 2; }
__host__ __device__ constexpr int gemm_smem_bytes(int block_n) {
    return gemm_stages(block_n) * gemm_stage_bytes(block_n) + 1024 /*align slack*/ + 384 /*barriers, tile ring*/;
}

// ---- PTX wrappers (mbarrier / TMA ones live in tma.cuh) -------------------------------------------
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (SWIZZLE_128B, descriptor version 1 for sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // version
    d |= (uint64_t)2 << 61;   // SWIZZLE_128B
    return d;
}

// Counter-based dropout hash shared by the GEMM epilogue and soswsod_dropout_mask: one splitmix64 round per
// PAIR of adjacent columns, 16 bits per element; keep iff bits >= p * 65536.
__host__ __device__ __forceinline__ uint32_t dropout_bits(unsigned long long seed, uint32_t m, uint32_t n) {
    unsigned long long x = seed + (((unsigned long long)m << 32) | (unsigned long long)(n >> 1)) * 0x9E3779B97F4A7C15ull;
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (n & 1u) ? (uint32_t)((x >> 16) & 0xFFFFu) : (uint32_t)(x & 0xFFFFu);
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
    float t = p * 65536.0f;
    if (t < 0.f) t = 0.f;
    if (t > 65535.f) t = 65535.f;
    return (uint32_t)t;
}

struct GemmEpilogue {
    const float* bias;
    int relu;
    const __nv_bfloat16* mask_src;
    long long ld_mask;
    float mask_scale;
    float dropout_p;
    unsigned long long dropout_seed;
};

template <int BLOCK_N, bool A_MN, bool B_MN, bool OUT_BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 void* __restrict__ dptr, long long ldd, int M, int N, int K, GemmEpilogue ep,
                 unsigned int* __restrict__ sched, int group_m) {
    constexpr int kStages = gemm_stages(BLOCK_N);
    constexpr int kABytes = kBlockM * kBlockK * 2;
    constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    constexpr int kStageBytes = kABytes + kBBytes;
    constexpr int kTmemCols = kNumAccStages * BLOCK_N;
    static_assert(kTmemCols == 256 || kTmemCols == 512, "TMEM columns must be a power of two");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + kStages * kStageBytes;
    // barriers: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]; then the TMEM base pointer
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + kNumAccStages + s); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * kStages + 2 * kNumAccStages);
    volatile uint32_t* tmem_ptr_generic =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    // tile scheduler: the producer publishes the tile index of every tile this CTA works on through a small ring
    // (full / empty mbarriers); with a scheduler workspace the next index comes from a global counter, so CTAs that
    // are slowed down (a co-resident NCCL kernel, a throttled SM) simply take fewer tiles.
    constexpr int kSchedSlots = 4;
    auto sfull_bar = [&](int s) { return tmem_ptr_addr + 8u + 8u * s; };
    auto sempty_bar = [&](int s) { return tmem_ptr_addr + 8u + 8u * (kSchedSlots + s); };
    const uint32_t sring_addr = tmem_ptr_addr + 8u + 8u * (2 * kSchedSlots);
    volatile int* sring = reinterpret_cast<volatile int*>(smem_raw + (sring_addr - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (M + kBlockM - 1) / kBlockM;
    const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (K + kBlockK - 1) / kBlockK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < kNumAccStages; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 128);
        }
        for (int s = 0; s < kSchedSlots; ++s) {
            mbar_init(sfull_bar(s), 1);
            mbar_init(sempty_bar(s), 5);   // MMA lane + one lane of each epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_generic;

    // Tile order: groups of `group_m` m-blocks, m fastest inside a group, so the ~148 tiles in flight form a
    // group_m x (148 / group_m) rectangle of the output: per wave the DRAM side reads group_m A row-blocks and
    // 148 / group_m B column-blocks.  group_m is chosen by the host to balance the two (gemm_group_m).
    const int kGroupM = group_m;
    auto tile_coords = [&](int t, int& mb, int& nb) {
        const int per_group = kGroupM * n_tiles;
        const int g = t / per_group;
        const int first_m = g * kGroupM;
        const int gsize = min(kGroupM, m_tiles - first_m);
        const int in = t - g * per_group;
        mb = first_m + in % gsize;
        nb = in / gsize;
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int ss = 0;
            uint32_t sphase = 0;
            int t = blockIdx.x;
            while (true) {
                mbar_wait(sempty_bar(ss), sphase ^ 1u);
                sring[ss] = t;
                mbar_arrive(sfull_bar(ss));
                if (++ss == kSchedSlots) {
                    ss = 0;
                    sphase ^= 1u;
                }
                if (t >= num_tiles) break;
                int mb, nb;
                tile_coords(t, mb, nb);
                const int m0 = mb * kBlockM, n0 = nb * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * kStageBytes;
                    const uint32_t sb = sa + kABytes;
                    mbar_expect_tx(full_bar(stage), kStageBytes);
                    const int k0 = kb * kBlockK;
                    if (!A_MN) {
                        tma_load_2d(sa, &tmap_a, full_bar(stage), k0, m0);
                    } else {
#pragma unroll
                        for (int c = 0; c < kBlockM / 64; ++c)
                            tma_load_2d(sa + c * (kBlockK * 128), &tmap_a, full_bar(stage), m0 + 64 * c, k0);
                    }
                    if (!B_MN) {
                        tma_load_2d(sb, &tmap_b, full_bar(stage), k0, n0);
                    } else {
#pragma unroll
                        for (int c = 0; c < BLOCK_N / 64; ++c)
                            tma_load_2d(sb + c * (kBlockK * 128), &tmap_b, full_bar(stage), n0 + 64 * c, k0);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                t = sched ? (int)(gridDim.x + atomicAdd(&sched[0], 1u)) : t + (int)gridDim.x;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                       ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                                       ((uint32_t)(kBlockM >> 4) << 24);
            // descriptor strides: K-major: SBO = 8 rows * 128 B; MN-major: SBO = 8 k-rows * 128 B, LBO = chunk stride
            constexpr uint32_t kLbo = kBlockK * 128, kSbo = 1024;
            constexpr uint32_t a_kstep = A_MN ? (kUmmaK * 128) : (kUmmaK * 2);
            constexpr uint32_t b_kstep = B_MN ? (kUmmaK * 128) : (kUmmaK * 2);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            int ss = 0;
            uint32_t sphase = 0;
            while (true) {
                mbar_wait(sfull_bar(ss), sphase);
                const int t = sring[ss];
                mbar_arrive(sempty_bar(ss));
                if (++ss == kSchedSlots) {
                    ss = 0;
                    sphase ^= 1u;
                }
                if (t >= num_tiles) break;
                mbar_wait(tempty_bar(as), aphase ^ 1u);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + stage * kStageBytes;
                    const uint32_t sb = sa + kABytes;
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        const uint64_t adesc = make_smem_desc(sa + k * a_kstep, A_MN ? kLbo : 0u, kSbo);
                        const uint64_t bdesc = make_smem_desc(sb + k * b_kstep, B_MN ? kLbo : 0u, kSbo);
                        tcgen05_mma_bf16(tmem_d, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    tcgen05_commit(empty_bar(stage));
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                tcgen05_commit(tfull_bar(as));
                if (++as == kNumAccStages) {
                    as = 0;
                    aphase ^= 1u;
                }
            }
        }
    } else {
        // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) ----
        const int q = warp & 3;
        const uint32_t thr = dropout_threshold(ep.dropout_p);
        const float drop_scale = ep.dropout_p > 0.f ? 1.0f / (1.0f - ep.dropout_p) : 1.0f;
        int as = 0;
        uint32_t aphase = 0;
        int ss = 0;
        uint32_t sphase = 0;
        while (true) {
            mbar_wait(sfull_bar(ss), sphase);
            const int t = sring[ss];
            __syncwarp();
            if (lane == 0) mbar_arrive(sempty_bar(ss));
            if (++ss == kSchedSlots) {
                ss = 0;
                sphase ^= 1u;
            }
            if (t >= num_tiles) break;
            int mb, nb;
            tile_coords(t, mb, nb);
            const int row = mb * kBlockM + q * 32 + lane;
            const int n0 = nb * BLOCK_N;
            mbar_wait(tfull_bar(as), aphase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N);
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr + c * 32, v);
                tmem_ld_wait();
                const int col0 = n0 + c * 32;
                if (row < M && col0 < N) {
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (ep.bias) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < N) f[j] += __ldg(ep.bias + col0 + j);
                    }
                    if (ep.relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                    }
                    if (ep.mask_src) {
                        const __nv_bfloat16* mrow = ep.mask_src + (size_t)row * ep.ld_mask + col0;
                        if (col0 + 32 <= N && ((reinterpret_cast<uintptr_t>(mrow) & 15) == 0)) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const uint4 u = __ldg(reinterpret_cast<const uint4*>(mrow) + g);
                                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    // bf16 > 0  <=>  sign bit clear and magnitude non-zero
                                    const uint32_t lo = w[i] & 0xFFFFu, hi = w[i] >> 16;
                                    f[g * 8 + 2 * i] *= (lo != 0u && lo < 0x8000u) ? ep.mask_scale : 0.f;
                                    f[g * 8 + 2 * i + 1] *= (hi != 0u && hi < 0x8000u) ? ep.mask_scale : 0.f;
                                }
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < N) f[j] *= (__bfloat162float(mrow[j]) > 0.f) ? ep.mask_scale : 0.f;
                        }
                    }
                    if (ep.dropout_p > 0.f) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            unsigned long long x = ep.dropout_seed +
                                                   (((unsigned long long)(uint32_t)row << 32) |
                                                    (unsigned long long)((uint32_t)(col0 + j) >> 1)) *
                                                       0x9E3779B97F4A7C15ull;
                            x ^= x >> 30;
                            x *= 0xBF58476D1CE4E5B9ull;
                            x ^= x >> 27;
                            x *= 0x94D049BB133111EBull;
                            x ^= x >> 31;
                            const uint32_t b0 = (uint32_t)(x & 0xFFFFu), b1 = (uint32_t)((x >> 16) & 0xFFFFu);
                            f[j] = (b0 >= thr) ? f[j] * drop_scale : 0.f;
                            f[j + 1] = (b1 >= thr) ? f[j + 1] * drop_scale : 0.f;
                        }
                    }
                    if (OUT_BF16) {
                        __nv_bfloat16* drow = reinterpret_cast<__nv_bfloat16*>(dptr) + (size_t)row * ldd + col0;
                        if (col0 + 32 <= N && ((reinterpret_cast<uintptr_t>(drow) & 15) == 0)) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                uint4 u;
                                __nv_bfloat162 p0 = __floats2bfloat162_rn(f[g * 8 + 0], f[g * 8 + 1]);
                                __nv_bfloat162 p1 = __floats2bfloat162_rn(f[g * 8 + 2], f[g * 8 + 3]);
                                __nv_bfloat162 p2 = __floats2bfloat162_rn(f[g * 8 + 4], f[g * 8 + 5]);
                                __nv_bfloat162 p3 = __floats2bfloat162_rn(f[g * 8 + 6], f[g * 8 + 7]);
                                u.x = *reinterpret_cast<uint32_t*>(&p0);
                                u.y = *reinterpret_cast<uint32_t*>(&p1);
                                u.z = *reinterpret_cast<uint32_t*>(&p2);
                                u.w = *reinterpret_cast<uint32_t*>(&p3);
                                reinterpret_cast<uint4*>(drow)[g] = u;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < N) drow[j] = __float2bfloat16_rn(f[j]);
                        }
                    } else {
                        float* drow = reinterpret_cast<float*>(dptr) + (size_t)row * ldd + col0;
                        if (col0 + 32 <= N && ((reinterpret_cast<uintptr_t>(drow) & 15) == 0)) {
#pragma unroll
                            for (int g = 0; g < 8; ++g)
                                reinterpret_cast<float4*>(drow)[g] =
                                    make_float4(f[g * 4 + 0], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < N) drow[j] = f[j];
                        }
                    }
                }
            }
            tcgen05_fence_before();
            mbar_arrive(tempty_bar(as));
            if (++as == kNumAccStages) {
                as = 0;
                aphase ^= 1u;
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
    // the last CTA to leave re-arms the scheduler workspace for the next launch on this stream
    if (sched && threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&sched[1], 1u) == gridDim.x - 1) {
            sched[0] = 0u;
            sched[1] = 0u;
            __threadfence();
        }
    }
}

// ---- host side -------------------------------------------------------------------------------------
int device_num_sms();

// A wave of `grid` tiles arranged as g m-blocks x (grid / g) n-blocks reads g * kBlockM + (grid / g) * BLOCK_N operand
// rows of K elements from DRAM (the rest hits L2): minimal for g = sqrt(grid * BLOCK_N / kBlockM), rounded to a power
// of two and clamped to the matrix.  SOSWSOD_GEMM_GROUP_M (read once) overrides it for experiments.
static int gemm_group_m(int m_tiles, int n_tiles, int block_n, int grid) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("SOSWSOD_GEMM_GROUP_M");
        forced = e ? atoi(e) : 0;
    }
    int g = forced;
    if (g <= 0) {
        const double ideal = sqrt((double)grid * block_n / kBlockM);
        g = 1;
        while (g * 2 <= ideal * 1.2) g *= 2;
    }
    if (g > m_tiles) g = m_tiles;
    return g < 1 ? 1 : g;
}

template <int BLOCK_N, bool A_MN, bool B_MN, bool OUT_BF16>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, void* d, long long ldd, int m, int n, int k,
                       const GemmEpilogue& ep, unsigned int* sched, cudaStream_t st) {
    auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN, OUT_BF16>;
    const int smem = gemm_smem_bytes(BLOCK_N);
    SOSWSOD_ENSURE_SMEM(kern, smem);
    const int m_tiles = (m + kBlockM - 1) / kBlockM, n_tiles = (n + BLOCK_N - 1) / BLOCK_N;
    const int tiles = m_tiles * n_tiles;
    const int grid = tiles < device_num_sms() ? tiles : device_num_sms();
    kern<<<grid, kGemmThreads, smem, st>>>(ta, tb, d, ldd, m, n, k, ep, sched, gemm_group_m(m_tiles, n_tiles, BLOCK_N, grid));
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

template <int BLOCK_N, bool OUT_BF16>
static int dispatch_major(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, void* d, long long ldd, int m,
                          int n, int k, const GemmEpilogue& ep, unsigned int* sched, cudaStream_t st) {
    if (!a_mn && !b_mn) return launch_gemm<BLOCK_N, false, false, OUT_BF16>(ta, tb, d, ldd, m, n, k, ep, sched, st);
    if (!a_mn && b_mn) return launch_gemm<BLOCK_N, false, true, OUT_BF16>(ta, tb, d, ldd, m, n, k, ep, sched, st);
    if (a_mn && !b_mn) return launch_gemm<BLOCK_N, true, false, OUT_BF16>(ta, tb, d, ldd, m, n, k, ep, sched, st);
    return launch_gemm<BLOCK_N, true, true, OUT_BF16>(ta, tb, d, ldd, m, n, k, ep, sched, st);
}

}  // namespace soswsod

using namespace soswsod;

extern "C" int soswsod_gemm_bf16(const void* a, long long lda, int a_mn_major, const void* b, long long ldb,
                                 int b_mn_major, void* d, long long ldd, int d_dtype, int m, int n, int k,
                                 const float* bias, int relu, const void* mask_src, long long ld_mask,
                                 float mask_scale, float dropout_p, unsigned long long dropout_seed,
                                 void* sched_workspace, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(a && b && d, "gemm_bf16: null pointer");
    SOSWSOD_CHECK_ARG(((uintptr_t)sched_workspace & 7) == 0, "gemm_bf16: scheduler workspace must be 8-byte aligned");
    unsigned int* sched = reinterpret_cast<unsigned int*>(sched_workspace);
    SOSWSOD_CHECK_ARG(m > 0 && n > 0 && k > 0, "gemm_bf16: bad shape m=%d n=%d k=%d", m, n, k);
    SOSWSOD_CHECK_ARG(d_dtype == SOSWSOD_DTYPE_F32 || d_dtype == SOSWSOD_DTYPE_BF16, "gemm_bf16: bad d_dtype");
    SOSWSOD_CHECK_ARG((lda % 8) == 0 && (ldb % 8) == 0, "gemm_bf16: lda/ldb must be multiples of 8 elements");
    SOSWSOD_CHECK_ARG(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0, "gemm_bf16: A/B must be 16-byte aligned");
    SOSWSOD_CHECK_ARG(lda >= (a_mn_major ? m : k) && ldb >= (b_mn_major ? n : k), "gemm_bf16: leading dimension too small");
    SOSWSOD_CHECK_ARG(ldd >= n, "gemm_bf16: ldd too small");
    SOSWSOD_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "gemm_bf16: dropout_p must be in [0,1)");
    SOSWSOD_CHECK_ARG(!mask_src || ld_mask >= n, "gemm_bf16: ld_mask too small");
    cudaStream_t st = (cudaStream_t)stream;
    // Narrow outputs use 128-wide tiles (more CTAs, less padding); everything else 256-wide.
    const int block_n = (n <= 128 || (n > 256 && n <= 384) || ((m + kBlockM - 1) / kBlockM) * ((n + 255) / 256) < 96) ? 128 : 256;
    CUtensorMap ta, tb;
    int rc;
    if (!a_mn_major) rc = make_tmap_2d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, m, k, lda, kBlockK, kBlockM, CU_TENSOR_MAP_SWIZZLE_128B);
    else rc = make_tmap_2d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, k, m, lda, 64, kBlockK, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (!b_mn_major) rc = make_tmap_2d(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, n, k, ldb, kBlockK, block_n, CU_TENSOR_MAP_SWIZZLE_128B);
    else rc = make_tmap_2d(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, k, n, ldb, 64, kBlockK, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    GemmEpilogue ep;
    ep.bias = bias;
    ep.relu = relu;
    ep.mask_src = reinterpret_cast<const __nv_bfloat16*>(mask_src);
    ep.ld_mask = ld_mask;
    ep.mask_scale = mask_scale;
    ep.dropout_p = dropout_p;
    ep.dropout_seed = dropout_seed;
    const bool obf = d_dtype == SOSWSOD_DTYPE_BF16;
    if (block_n == 256) {
        if (obf) return dispatch_major<256, true>(a_mn_major, b_mn_major, ta, tb, d, ldd, m, n, k, ep, sched, st);
        return dispatch_major<256, false>(a_mn_major, b_mn_major, ta, tb, d, ldd, m, n, k, ep, sched, st);
    }
    if (obf) return dispatch_major<128, true>(a_mn_major, b_mn_major, ta, tb, d, ldd, m, n, k, ep, sched, st);
    return dispatch_major<128, false>(a_mn_major, b_mn_major, ta, tb, d, ldd, m, n, k, ep, sched, st);
}

namespace soswsod {
__global__ void dropout_mask_kernel(unsigned char* __restrict__ mask, int m, int n, uint32_t thr,
                                    unsigned long long seed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)m * n) return;
    const uint32_t r = (uint32_t)(i / n), c = (uint32_t)(i % n);
    mask[i] = dropout_bits(seed, r, c) >= thr ? 1 : 0;
}
}  // namespace soswsod

extern "C" int soswsod_dropout_mask(unsigned char* mask, int m, int n, float dropout_p, unsigned long long seed,
                                    soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(mask && m > 0 && n > 0, "dropout_mask: bad arguments");
    SOSWSOD_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "dropout_mask: dropout_p must be in [0,1)");
    const long long total = (long long)m * n;
    dropout_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        mask, m, n, dropout_p > 0.f ? dropout_threshold(dropout_p) : 0u, seed);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}
