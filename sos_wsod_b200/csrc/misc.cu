// Error plumbing of the C ABI plus the small HBM-bound helpers around the GEMMs: fp32 -> bf16 cast (with
// optional per-column scale and transposed copy), bf16 transpose, column sums (bias gradients).
#include <stdarg.h>

#include "common.cuh"

namespace soswsod {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

constexpr int kTile = 64;

// in fp32 [rows, ld_in] -> out bf16 [rows, ld_out] and/or out_t bf16 [cols, ld_out_t]; 64x64 tiles, 256 threads.
__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ in, long long ld_in, int rows, int cols,
                     const float* __restrict__ col_scale, __nv_bfloat16* __restrict__ out, long long ld_out,
                     __nv_bfloat16* __restrict__ out_t, long long ld_out_t,
                     const __nv_bfloat16* __restrict__ mask_src, long long ld_mask, float mask_scale) {
    __shared__ __nv_bfloat16 tile[kTile][kTile + 2];
    const int r0 = blockIdx.y * kTile, c0 = blockIdx.x * kTile;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
    const int c = c0 + tx;
    const float s = (col_scale && c < cols) ? col_scale[c] : 1.f;
    for (int i = ty; i < kTile; i += 4) {
        const int r = r0 + i;
        __nv_bfloat16 v = __float2bfloat16_rn(0.f);
        if (r < rows && c < cols) {
            float f = in[(size_t)r * ld_in + c] * s;
            // backward of ReLU (+ dropout): pass the gradient where the saved activation is positive
            if (mask_src) f *= (__bfloat162float(mask_src[(size_t)r * ld_mask + c]) > 0.f) ? mask_scale : 0.f;
            v = __float2bfloat16_rn(f);
            if (out) out[(size_t)r * ld_out + c] = v;
        }
        tile[i][tx] = v;
    }
    if (!out_t) return;
    __syncthreads();
    const int r = r0 + tx;
    for (int i = ty; i < kTile; i += 4) {
        const int cc = c0 + i;
        if (r < rows && cc < cols) out_t[(size_t)cc * ld_out_t + r] = tile[tx][i];
    }
}

__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long ld_in, int rows, int cols,
                      __nv_bfloat16* __restrict__ out_t, long long ld_out_t) {
    __shared__ __nv_bfloat16 tile[kTile][kTile + 2];
    const int r0 = blockIdx.y * kTile, c0 = blockIdx.x * kTile;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int i = ty; i < kTile; i += 4) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? in[(size_t)r * ld_in + c] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    const int r = r0 + tx;
    for (int i = ty; i < kTile; i += 4) {
        const int cc = c0 + i;
        if (r < rows && cc < cols) out_t[(size_t)cc * ld_out_t + r] = tile[tx][i];
    }
}

// out[c] = sum_r in[r, c]; block = 32 columns x 32 row lanes, fixed summation order (deterministic).
template <typename T>
__global__ void __launch_bounds__(1024)
colsum_kernel(const T* __restrict__ in, long long ld_in, int rows, int cols, float* __restrict__ out) {
    __shared__ float part[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float acc = 0.f;
    if (c < cols) {
        for (int r = ty; r < rows; r += 32) {
            if (sizeof(T) == 2)
                acc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&in[(size_t)r * ld_in + c]));
            else
                acc += *reinterpret_cast<const float*>(&in[(size_t)r * ld_in + c]);
        }
    }
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < cols) {
        float s = 0.f;
        for (int i = 0; i < 32; ++i) s += part[i][tx];
        out[c] = s;
    }
}

// Two-pass column sum for wide matrices: pass 1 streams the matrix once with 16-byte loads (a thread owns VEC
// adjacent columns of one row chunk) and writes per-chunk partial sums to the workspace [chunks, cols]; pass 2 adds
// the chunks of a column in chunk order.  Deterministic, no atomics.
template <typename T, int VEC>
__global__ void __launch_bounds__(128)
colsum_partial_kernel(const T* __restrict__ in, long long ld_in, int rows, int cols, int rows_per_chunk,
                      float* __restrict__ part) {
    const int c = (blockIdx.x * 128 + threadIdx.x) * VEC;
    if (c >= cols) return;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    const T* p = in + (size_t)r0 * ld_in + c;
#pragma unroll 4
    for (int r = r0; r < r1; ++r, p += ld_in) {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        if (sizeof(T) == 2) {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc[2 * k] += __uint_as_float(w[k] << 16);
                acc[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
            }
        } else {
            acc[0] += __uint_as_float(v.x);
            acc[1] += __uint_as_float(v.y);
            acc[2] += __uint_as_float(v.z);
            acc[3] += __uint_as_float(v.w);
        }
    }
    float* dst = part + (size_t)blockIdx.y * cols + c;
#pragma unroll
    for (int k = 0; k < VEC; ++k) dst[k] = acc[k];
}

__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ part, int chunks, int cols, float* __restrict__ out) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= cols) return;
    float s = 0.f;
    for (int k = 0; k < chunks; ++k) s += part[(size_t)k * cols + c];
    out[c] = s;
}

static int colsum_chunks(int rows, int cols, int vec) {
    const int col_blocks = (cols / vec + 127) / 128;
    int chunks = (2 * 148 + col_blocks - 1) / col_blocks;
    if (chunks > (rows + 15) / 16) chunks = (rows + 15) / 16;   // at least 16 rows per chunk
    return chunks < 1 ? 1 : chunks;
}

// SGD with momentum and weight decay (torch.optim.SGD semantics, as built by uwsod/detectron2/solver/build.py:
// g = grad*grad_scale + wd*p ; buf = momentum*buf + g ; p -= lr*buf), fused with the refresh of the bf16 GEMM
// operand copy of the parameter.
__global__ void sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n,
                                float lr, float momentum, float wd, float gscale, __nv_bfloat16* __restrict__ p_bf16) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float pv = p[i];
    const float gv = g[i] * gscale + wd * pv;
    const float b = momentum * buf[i] + gv;
    buf[i] = b;
    const float np = pv - lr * b;
    p[i] = np;
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(np);
}

// The same update for up to SOSWSOD_SGD_MAX_TENSORS parameter tensors (or shards of them) in ONE launch: the table of
// tensors travels in the kernel parameters, a block owns 4096 consecutive elements of one tensor, 128-bit accesses
// where the tensor's pointers allow.  Sinks: the bf16 GEMM-operand copy and / or an fp32 copy of the new value.
constexpr int kSgdBlockElems = 4096;

struct SgdBatch {
    soswsod_sgd_tensor t[SOSWSOD_SGD_MAX_TENSORS];
    int block_start[SOSWSOD_SGD_MAX_TENSORS + 1];
    int count;
    float momentum, gscale;
};

__global__ void __launch_bounds__(256)
sgd_multi_kernel(const __grid_constant__ SgdBatch b) {
    int ti = 0;
    while (ti + 1 < b.count && (int)blockIdx.x >= b.block_start[ti + 1]) ++ti;
    const soswsod_sgd_tensor& t = b.t[ti];
    const long long e0 = (long long)((int)blockIdx.x - b.block_start[ti]) * kSgdBlockElems;
    const long long e1 = e0 + kSgdBlockElems < t.n ? e0 + kSgdBlockElems : t.n;
    float* __restrict__ p = t.param;
    const float* __restrict__ g = t.grad;
    float* __restrict__ mb = t.momentum_buf;
    __nv_bfloat16* __restrict__ ob = reinterpret_cast<__nv_bfloat16*>(t.out_bf16);
    float* __restrict__ of = t.out_f32;
    const float lr = t.lr, wd = t.weight_decay, mom = b.momentum, gs = b.gscale;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(mb) |
                       reinterpret_cast<uintptr_t>(of)) & 15) == 0 && (reinterpret_cast<uintptr_t>(ob) & 7) == 0;
    auto upd = [&](float pv, float gv, float bv, float& nb) {
        nb = mom * bv + (gv * gs + wd * pv);
        return pv - lr * nb;
    };
    long long i = e0 + 4LL * threadIdx.x;
    if (vec) {
        for (; i + 4 <= e1; i += 4 * 256) {
            const float4 pv = *reinterpret_cast<const float4*>(p + i);
            const float4 gv = __ldcs(reinterpret_cast<const float4*>(g + i));      // the gradient is read once
            const float4 bv = *reinterpret_cast<const float4*>(mb + i);
            float4 nb, np;
            np.x = upd(pv.x, gv.x, bv.x, nb.x);
            np.y = upd(pv.y, gv.y, bv.y, nb.y);
            np.z = upd(pv.z, gv.z, bv.z, nb.z);
            np.w = upd(pv.w, gv.w, bv.w, nb.w);
            *reinterpret_cast<float4*>(mb + i) = nb;
            *reinterpret_cast<float4*>(p + i) = np;
            if (ob) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(np.x, np.y), hi = __floats2bfloat162_rn(np.z, np.w);
                uint2 u;
                u.x = *reinterpret_cast<const uint32_t*>(&lo);
                u.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(ob + i) = u;
            }
            if (of) *reinterpret_cast<float4*>(of + i) = np;
        }
        // tail (n % 4) of the tensor: the lanes past the last full vector
        if (i < e1 && i + 4 > e1) {
            for (long long j = i; j < e1; ++j) {
                float nb;
                const float np = upd(p[j], g[j], mb[j], nb);
                mb[j] = nb;
                p[j] = np;
                if (ob) ob[j] = __float2bfloat16_rn(np);
                if (of) of[j] = np;
            }
        }
    } else {
        for (long long j = e0 + threadIdx.x; j < e1; j += 256) {
            float nb;
            const float np = upd(p[j], g[j], mb[j], nb);
            mb[j] = nb;
            p[j] = np;
            if (ob) ob[j] = __float2bfloat16_rn(np);
            if (of) of[j] = np;
        }
    }
}

// The data-parallel update as ONE kernel over NVSwitch multicast memory (the fused compute + collective of this path):
// every rank owns a block of rows of the big weight matrices.  For its rows it
//   * reads the SUM over all ranks of the gradient with multimem.ld_reduce (the switch adds the ranks' copies: the
//     reduce-scatter, with no ring and no staging buffer),
//   * applies SGD + momentum + weight decay to its rows of the fp32 master and momentum buffer (local memory),
//   * writes the refreshed bf16 GEMM-operand rows with multimem.st, which lands them in EVERY rank's operand matrix (the
//     all-gather).
// grad_mc / out_bf16_mc are multicast addresses of symmetric buffers (same offset on every rank); the caller brackets
// the launch with cross-rank barriers (all gradients written before, all operand rows landed after).
static int device_num_sms_misc() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
        else sms = 148;
    }
    return sms;
}

constexpr int kNvlsThreads = 512;
constexpr int kNvlsPerThread = 16;                              // elements per thread and iteration: 4 x 16-byte ld_reduce in flight
constexpr int kNvlsChunk = kNvlsThreads * kNvlsPerThread;       // 8192 elements

struct SgdNvlsBatch {
    soswsod_sgd_nvls_tensor t[SOSWSOD_SGD_NVLS_MAX_TENSORS];
    long long chunk_start[SOSWSOD_SGD_NVLS_MAX_TENSORS + 1];
    int count;
    float momentum, gscale;
};

// Persistent grid: the switch serves a bounded number of outstanding multimem requests well (NCCL / torch drive NVLS with
// a few dozen CTAs); every CTA walks the chunk list with stride gridDim.x, every thread keeps four 16-byte reductions in
// flight per iteration.
__global__ void __launch_bounds__(kNvlsThreads)
sgd_nvls_kernel(const __grid_constant__ SgdNvlsBatch b) {
    const long long total = b.chunk_start[b.count];
    for (long long c = blockIdx.x; c < total; c += gridDim.x) {
        int ti = 0;
        while (ti + 1 < b.count && c >= b.chunk_start[ti + 1]) ++ti;
        const soswsod_sgd_nvls_tensor& t = b.t[ti];
        const long long i = (c - b.chunk_start[ti]) * kNvlsChunk + (long long)threadIdx.x * 4;
        float* __restrict__ p = t.param;
        float* __restrict__ mb = t.momentum_buf;
        const float* g_mc = t.grad_mc;
        __nv_bfloat16* ob_mc = reinterpret_cast<__nv_bfloat16*>(t.out_bf16_mc);
        const float lr = t.lr, wd = t.weight_decay, mom = b.momentum, gs = b.gscale;
        // four groups of 4 consecutive elements, kNvlsThreads * 4 apart: a warp's 32 lanes read 512 contiguous bytes
        float4 g[4];
        bool live[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long e = i + (long long)k * kNvlsThreads * 4;
            live[k] = e + 4 <= t.n;
            if (live[k])
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(g[k].x), "=f"(g[k].y), "=f"(g[k].z), "=f"(g[k].w) : "l"(g_mc + e) : "memory");
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!live[k]) continue;
            const long long e = i + (long long)k * kNvlsThreads * 4;
            const float4 pv = *reinterpret_cast<const float4*>(p + e);
            const float4 bv = *reinterpret_cast<const float4*>(mb + e);
            float4 nb, np;
            nb.x = mom * bv.x + (g[k].x * gs + wd * pv.x); np.x = pv.x - lr * nb.x;
            nb.y = mom * bv.y + (g[k].y * gs + wd * pv.y); np.y = pv.y - lr * nb.y;
            nb.z = mom * bv.z + (g[k].z * gs + wd * pv.z); np.z = pv.z - lr * nb.z;
            nb.w = mom * bv.w + (g[k].w * gs + wd * pv.w); np.w = pv.w - lr * nb.w;
            *reinterpret_cast<float4*>(mb + e) = nb;
            *reinterpret_cast<float4*>(p + e) = np;
            const __nv_bfloat162 q0 = __floats2bfloat162_rn(np.x, np.y), q1 = __floats2bfloat162_rn(np.z, np.w);
            asm volatile("multimem.st.relaxed.sys.global.v2.bf16x2 [%0], {%1,%2};" ::"l"(ob_mc + e),
                         "r"(*reinterpret_cast<const uint32_t*>(&q0)), "r"(*reinterpret_cast<const uint32_t*>(&q1)) : "memory");
        }
    }
    __threadfence_system();   // this thread's multicast stores are performed system-wide before the kernel retires
}

}  // namespace soswsod

using namespace soswsod;

extern "C" int soswsod_sgd_nvls(const soswsod_sgd_nvls_tensor* tensors, int count, float momentum, float grad_scale,
                                int max_ctas, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(tensors && count > 0 && count <= SOSWSOD_SGD_NVLS_MAX_TENSORS, "sgd_nvls: 1..%d tensors per call",
                      SOSWSOD_SGD_NVLS_MAX_TENSORS);
    SgdNvlsBatch b;
    b.count = count;
    b.momentum = momentum;
    b.gscale = grad_scale;
    long long chunks = 0;
    for (int i = 0; i < count; ++i) {
        const soswsod_sgd_nvls_tensor& t = tensors[i];
        SOSWSOD_CHECK_ARG(t.param && t.grad_mc && t.momentum_buf && t.out_bf16_mc && t.n > 0, "sgd_nvls: tensor %d: null pointer or empty", i);
        SOSWSOD_CHECK_ARG(t.n % 4 == 0, "sgd_nvls: tensor %d: %lld elements, need a multiple of 4", i, t.n);
        SOSWSOD_CHECK_ARG(((reinterpret_cast<uintptr_t>(t.param) | reinterpret_cast<uintptr_t>(t.grad_mc) |
                            reinterpret_cast<uintptr_t>(t.momentum_buf)) & 15) == 0 &&
                              (reinterpret_cast<uintptr_t>(t.out_bf16_mc) & 7) == 0,
                          "sgd_nvls: tensor %d: fp32 pointers must be 16-byte, the bf16 operand 8-byte aligned", i);
        b.t[i] = t;
        b.chunk_start[i] = chunks;
        chunks += (t.n + kNvlsChunk - 1) / kNvlsChunk;
    }
    b.chunk_start[count] = chunks;
    if (max_ctas <= 0) max_ctas = device_num_sms_misc();
    const long long grid = chunks < max_ctas ? chunks : max_ctas;
    sgd_nvls_kernel<<<(unsigned)grid, kNvlsThreads, 0, (cudaStream_t)stream>>>(b);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

extern "C" int soswsod_sgd_multi(const soswsod_sgd_tensor* tensors, int count, float momentum, float grad_scale,
                                 soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(tensors && count > 0 && count <= SOSWSOD_SGD_MAX_TENSORS, "sgd_multi: 1..%d tensors per call",
                      SOSWSOD_SGD_MAX_TENSORS);
    SgdBatch b;
    b.count = count;
    b.momentum = momentum;
    b.gscale = grad_scale;
    long long blocks = 0;
    for (int i = 0; i < count; ++i) {
        SOSWSOD_CHECK_ARG(tensors[i].param && tensors[i].grad && tensors[i].momentum_buf && tensors[i].n > 0,
                          "sgd_multi: tensor %d: null pointer or empty", i);
        b.t[i] = tensors[i];
        b.block_start[i] = (int)blocks;
        blocks += (tensors[i].n + kSgdBlockElems - 1) / kSgdBlockElems;
        SOSWSOD_CHECK_ARG(blocks < (1LL << 30), "sgd_multi: too many elements for one launch");
    }
    b.block_start[count] = (int)blocks;
    sgd_multi_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(b);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

extern "C" int soswsod_sgd_step(float* param, const float* grad, float* momentum_buf, long long n, float lr,
                                float momentum, float weight_decay, float grad_scale, void* param_bf16,
                                soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(param && grad && momentum_buf && n > 0, "sgd_step: bad arguments");
    sgd_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, momentum_buf, n, lr, momentum,
                                                                                weight_decay, grad_scale,
                                                                                (__nv_bfloat16*)param_bf16);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

extern "C" int soswsod_abi_version(void) { return SOSWSOD_ABI_VERSION; }
extern "C" const char* soswsod_last_error(void) { return g_err; }

extern "C" int soswsod_cast_f32_bf16(const float* in, long long ld_in, int rows, int cols, const float* col_scale,
                                     void* out, long long ld_out, void* out_t, long long ld_out_t,
                                     const void* mask_src, long long ld_mask, float mask_scale,
                                     soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(in && (out || out_t), "cast_f32_bf16: null pointer");
    SOSWSOD_CHECK_ARG(rows > 0 && cols > 0 && ld_in >= cols, "cast_f32_bf16: bad shape");
    SOSWSOD_CHECK_ARG(!out || ld_out >= cols, "cast_f32_bf16: ld_out too small");
    SOSWSOD_CHECK_ARG(!out_t || ld_out_t >= rows, "cast_f32_bf16: ld_out_t too small");
    SOSWSOD_CHECK_ARG(!mask_src || ld_mask >= cols, "cast_f32_bf16: ld_mask too small");
    dim3 grid((cols + kTile - 1) / kTile, (rows + kTile - 1) / kTile);
    cast_f32_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, ld_in, rows, cols, col_scale,
                                                                 (__nv_bfloat16*)out, ld_out, (__nv_bfloat16*)out_t,
                                                                 ld_out_t, (const __nv_bfloat16*)mask_src, ld_mask, mask_scale);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

extern "C" int soswsod_transpose_bf16(const void* in, long long ld_in, int rows, int cols, void* out_t,
                                      long long ld_out_t, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(in && out_t, "transpose_bf16: null pointer");
    SOSWSOD_CHECK_ARG(rows > 0 && cols > 0 && ld_in >= cols && ld_out_t >= rows, "transpose_bf16: bad shape");
    dim3 grid((cols + kTile - 1) / kTile, (rows + kTile - 1) / kTile);
    transpose_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, ld_in, rows, cols,
                                                                  (__nv_bfloat16*)out_t, ld_out_t);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}

extern "C" size_t soswsod_colsum_workspace_bytes(int rows, int cols, int in_dtype) {
    const int vec = in_dtype == SOSWSOD_DTYPE_BF16 ? 8 : 4;
    if (rows <= 0 || cols <= 0 || cols % vec) return 0;
    return (size_t)colsum_chunks(rows, cols, vec) * cols * 4;
}

extern "C" int soswsod_colsum(const void* in, int in_dtype, long long ld_in, int rows, int cols, float* out,
                              void* workspace, size_t workspace_bytes, soswsod_stream_t stream) {
    SOSWSOD_CHECK_ARG(in && out, "colsum: null pointer");
    SOSWSOD_CHECK_ARG(rows > 0 && cols > 0 && ld_in >= cols, "colsum: bad shape");
    SOSWSOD_CHECK_ARG(in_dtype == SOSWSOD_DTYPE_F32 || in_dtype == SOSWSOD_DTYPE_BF16, "colsum: bad dtype");
    const int vec = in_dtype == SOSWSOD_DTYPE_BF16 ? 8 : 4;
    const int esz = in_dtype == SOSWSOD_DTYPE_BF16 ? 2 : 4;
    if (workspace && cols % vec == 0 && ((ld_in * esz) & 15) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(workspace) & 15) == 0) {
        const int chunks = colsum_chunks(rows, cols, vec);
        if (workspace_bytes >= (size_t)chunks * cols * 4 && chunks > 1) {
            const int rows_per_chunk = (rows + chunks - 1) / chunks;
            const dim3 grid((cols / vec + 127) / 128, chunks);
            if (in_dtype == SOSWSOD_DTYPE_BF16)
                colsum_partial_kernel<__nv_bfloat16, 8><<<grid, 128, 0, (cudaStream_t)stream>>>(
                    (const __nv_bfloat16*)in, ld_in, rows, cols, rows_per_chunk, (float*)workspace);
            else
                colsum_partial_kernel<float, 4><<<grid, 128, 0, (cudaStream_t)stream>>>((const float*)in, ld_in, rows, cols,
                                                                                     rows_per_chunk, (float*)workspace);
            SOSWSOD_CHECK_LAUNCH();
            colsum_final_kernel<<<(cols + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float*)workspace, chunks, cols, out);
            SOSWSOD_CHECK_LAUNCH();
            return SOSWSOD_OK;
        }
    }
    const int grid = (cols + 31) / 32;
    if (in_dtype == SOSWSOD_DTYPE_BF16)
        colsum_kernel<__nv_bfloat16><<<grid, 1024, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, ld_in, rows, cols, out);
    else
        colsum_kernel<float><<<grid, 1024, 0, (cudaStream_t)stream>>>((const float*)in, ld_in, rows, cols, out);
    SOSWSOD_CHECK_LAUNCH();
    return SOSWSOD_OK;
}
