// skrample_b200 - noise generation kernels for sm_100a (B200).
//
// Replaces the ATen op chains of skrample/pytorch/noise.py (reference lines cited per kernel):
//   skr_noise_fill      Random / Offset: counter-based Philox4x32-10 normals written straight in the
//                       storage dtype, 128-bit stores, optional running sum / sum-of-squares.
//   skr_noise_pyramid   Pyramid: base + full-resolution level + bilinear-upsampled coarse levels, evaluated
//                       per element from Philox streams (or supplied buffers), two passes: moments, then
//                       regenerate + normalise + write - the N-sized intermediate is never stored.
//   skr_noise_brownian  Brownian: increment of a seed-keyed Brownian path over (t0, t1), a stateless Philox
//                       bridge tree evaluated per element (replaces torchsde's interval tree on CUDA).
//   skr_noise_moments   sum / sum^2 of a tensor (for std), warp-shuffle + block reduction, fp64 partials.
//   skr_noise_scale     out = in * scale (in place allowed), storage-dtype aware.
//   skr_colored_shape   in-place spectral shaping of an rfftn half-spectrum: multiply each complex bin by
//                       clamp(radial_frequency, eps)^(-exponent/2); the radial grid is computed from indices.
//
// Philox streams are keyed (seed, stream id) with the counter = element index / 4, so any batch sharding
// over GPUs reproduces the same values for the same (seed, stream, index).

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/skrample_b200.h"
#include "common.cuh"
#include "machine.cuh"

namespace skr {

extern int fail(int code, const char* fmt, ...);
extern void count_launch(int kind);
extern int sm_count_or(int fallback);
extern int env_int(const char* name, int fallback);

// ---------------------------------------------------------------------------------------------------------
// store helpers

__device__ __forceinline__ void store4(void* out, int dtype, int64_t first, const float (&v)[4]) {
    switch (dtype) {
        case SKR_F32: *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + first) = make_float4(v[0], v[1], v[2], v[3]); break;
        case SKR_BF16: {
            const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + first) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
        } break;
        case SKR_F16: {
            const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + first) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
        } break;
        default: {
            double* p = reinterpret_cast<double*>(out) + first;
            *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
            *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
        } break;
    }
}

__device__ __forceinline__ void store1(void* out, int dtype, int64_t e, float v) {
    switch (dtype) {
        case SKR_F32: reinterpret_cast<float*>(out)[e] = v; break;
        case SKR_BF16: reinterpret_cast<__nv_bfloat16*>(out)[e] = __float2bfloat16_rn(v); break;
        case SKR_F16: reinterpret_cast<__half*>(out)[e] = __float2half_rn(v); break;
        default: reinterpret_cast<double*>(out)[e] = (double)v; break;
    }
}

__device__ __forceinline__ float load1(const void* in, int dtype, int64_t e) {
    switch (dtype) {
        case SKR_F32: return reinterpret_cast<const float*>(in)[e];
        case SKR_BF16: return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(in)[e]);
        case SKR_F16: return __half2float(reinterpret_cast<const __half*>(in)[e]);
        default: return (float)reinterpret_cast<const double*>(in)[e];
    }
}

// block-wide sum of two doubles; result valid in thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b) {
    __shared__ double sa[32], sb[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sa[warp] = a; sb[warp] = b; }
    __syncthreads();
    if (warp == 0) {
        const int n = (blockDim.x + 31) >> 5;
        a = lane < n ? sa[lane] : 0.0;
        b = lane < n ? sb[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
        }
    }
}

// Grid-wide sum of two doubles with a FIXED summation order, so the std that normalises Pyramid / Colored noise and
// the RKMoire error ratio are bit-reproducible from run to run (atomicAdd on doubles would add the block sums in
// arrival order).  `buf` is the caller's accumulator, SKR_MOMENTS_DOUBLES doubles, zeroed before first use:
//   buf[0..1]  the two sums (added to what is there: several launches may accumulate into one buffer, in stream order)
//   buf[2]     arrival counter (bit pattern of an unsigned integer), left at zero again
//   buf[4 + 2b], buf[5 + 2b]  partial sums of block b
// Every block stores its partials; the block that arrives last adds them up in block order (each lane a fixed strided
// subset, then a fixed shuffle tree) and publishes the totals.  gridDim.x <= SKR_MOMENT_BLOCKS.
//   buf[3]     (resident kernels only) generation counter: bumped once the totals of a launch are published
template <bool SIGNAL = false>
__device__ __forceinline__ void publish_sums(double a, double b, double* buf) {
    __shared__ bool last;
    block_sum2(a, b);
    if (threadIdx.x == 0) {
        buf[4 + 2 * blockIdx.x] = a;
        buf[5 + 2 * blockIdx.x] = b;
        __threadfence();
        const unsigned int ticket = atomicAdd(reinterpret_cast<unsigned int*>(buf + 2), 1u);
        last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        __threadfence();
        double ta = 0.0, tb = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += 32) {
            ta += __ldcg(buf + 4 + 2 * i);
            tb += __ldcg(buf + 5 + 2 * i);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ta += __shfl_down_sync(0xffffffffu, ta, o);
            tb += __shfl_down_sync(0xffffffffu, tb, o);
        }
        if (threadIdx.x == 0) {
            buf[0] += ta;
            buf[1] += tb;
            *reinterpret_cast<unsigned int*>(buf + 2) = 0u;
            if constexpr (SIGNAL) {
                __threadfence();
                atomicAdd(reinterpret_cast<unsigned int*>(buf + 3), 1u);
            }
        }
    }
}

__device__ __forceinline__ unsigned int load_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// 128-bit loads / stores of VEC consecutive elements of a storage type, as floats
template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
    static __device__ __forceinline__ float get(const float* p) { return *p; }
    static __device__ __forceinline__ void put(float* p, float v) { *p = v; }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 q = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    static __device__ __forceinline__ float get(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void put(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};
template <> struct Vec<__half> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
        const uint4 q = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__half* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    static __device__ __forceinline__ float get(const __half* p) { return __half2float(*p); }
    static __device__ __forceinline__ void put(__half* p, float v) { *p = __float2half_rn(v); }
};

// ---------------------------------------------------------------------------------------------------------
// Random / Offset fill   (reference: noise.py:36-42,73-74,104-113)

struct FillParams {
    void* out;
    int64_t numel;
    uint64_t seed, stream, offset_stream;
    float offset_scale;  // strength^2; 0 disables the offset term
    int32_t dtype, ndim, aligned;
    int64_t shape[SKR_MAX_DIMS];
    int32_t keep[SKR_MAX_DIMS];  // 1: axis indexes the offset tensor, 0: broadcast
    double* moments;             // optional [sum, sum^2] accumulators
    int64_t inner;               // > 0: the kept axes are a leading prefix, offset index = element / inner
};

__device__ __forceinline__ int64_t reduced_index(const FillParams& p, int64_t e) {
    // linear index into the offset tensor (shape = shape[d] if keep[d] else 1), row-major
    int64_t idx[SKR_MAX_DIMS];
    for (int d = p.ndim - 1; d >= 0; --d) { idx[d] = e % p.shape[d]; e /= p.shape[d]; }
    int64_t r = 0;
    for (int d = 0; d < p.ndim; ++d) if (p.keep[d]) r = r * p.shape[d] + idx[d];
    return r;
}

// four consecutive elements in storage type T
template <typename T> __device__ __forceinline__ void put4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void put4<float>(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
template <> __device__ __forceinline__ void put4<double>(double* p, const float (&v)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}
template <> __device__ __forceinline__ void put4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
template <> __device__ __forceinline__ void put4<__half>(__half* p, const float (&v)[4]) {
    const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
template <typename T> __device__ __forceinline__ void put1(T* p, float v) { *p = (T)v; }
template <> __device__ __forceinline__ void put1<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void put1<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// T (the storage type), OFFSET and MOMENTS are compile-time so that the plain Random fill carries neither a dtype
// dispatch, nor the offset lookup, nor the sums: its loop is Philox + Box-Muller + one 128-bit store.
template <typename T, bool OFFSET, bool MOMENTS>
__global__ void __launch_bounds__(256) fill_kernel(const __grid_constant__ FillParams p) {
    const Philox ph(p.seed);
    T* const out = reinterpret_cast<T*>(p.out);
    double s1 = 0.0, s2 = 0.0;
    const int64_t groups = (p.numel + 3) >> 2;
    const int64_t whole = p.aligned ? p.numel >> 2 : 0;  // groups stored with one vector access
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        float z[4];
        normal4(ph((uint64_t)g, p.stream), z);
        const int64_t first = g << 2;
        if constexpr (OFFSET) {
            const int64_t row = p.inner > 0 ? first / p.inner : -1;
            if (row >= 0 && first + 3 < (row + 1) * p.inner) {
                // the usual case (offsets along leading axes): the four elements share one offset draw
                // (a product and a sum, individually rounded: the in-step draw of machine.cuh computes the same bits)
                const float shift = __fmul_rn(normal_at(ph, (uint64_t)row, p.offset_stream), p.offset_scale);
#pragma unroll
                for (int j = 0; j < 4; ++j) z[j] = __fadd_rn(z[j], shift);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (first + j < p.numel)
                        z[j] = __fadd_rn(z[j], __fmul_rn(normal_at(ph, (uint64_t)reduced_index(p, first + j), p.offset_stream), p.offset_scale));
                }
            }
        }
        if (g < whole) {
            put4<T>(out + first, z);
            if constexpr (MOMENTS) {  // four values in fp32 (exact enough: |z| < 7), then one fp64 add per sum
                s1 += (double)((z[0] + z[1]) + (z[2] + z[3]));
                s2 += (double)((z[0] * z[0] + z[1] * z[1]) + (z[2] * z[2] + z[3] * z[3]));
            }
        } else {
            for (int j = 0; j < 4 && first + j < p.numel; ++j) {
                put1<T>(out + first + j, z[j]);
                if constexpr (MOMENTS) { s1 += z[j]; s2 += (double)z[j] * z[j]; }
            }
        }
    }
    if constexpr (MOMENTS) publish_sums(s1, s2, p.moments);
}

template <typename T>
static void launch_fill(const FillParams& p, bool shifted, bool moments, unsigned grid, cudaStream_t s) {
    if (moments) {
        if (shifted) fill_kernel<T, true, true><<<grid, 256, 0, s>>>(p);
        else fill_kernel<T, false, true><<<grid, 256, 0, s>>>(p);
    } else {
        if (shifted) fill_kernel<T, true, false><<<grid, 256, 0, s>>>(p);
        else fill_kernel<T, false, false><<<grid, 256, 0, s>>>(p);
    }
}

// Batched Random fill: one launch for every batch item, each with its own (seed, stream) - the same values
// skr_noise_fill writes item by item (reference: BatchTensorNoise.generate, noise.py:445-446).
struct BatchFillParams {
    void* out;
    int64_t numel;
    int32_t dtype, aligned;
    KPhilox keys;
};

template <typename T>
__global__ void __launch_bounds__(256) batch_fill_kernel(const __grid_constant__ BatchFillParams p) {
    T* const out = reinterpret_cast<T*>(p.out);
    const int64_t groups = (p.numel + 3) >> 2;
    const int64_t whole = p.aligned ? p.numel >> 2 : 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t first = g << 2;
        float z[4];
        draw_normals<float, 4>(p.keys, first, p.numel, z);
        if (g < whole) put4<T>(out + first, z);
        else for (int j = 0; j < 4 && first + j < p.numel; ++j) put1<T>(out + first + j, z[j]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// moments / scale   (reference: noise.py:207,365,401-403 - Tensor.std())

struct MomentParams {
    const void* in;
    int64_t numel;
    int32_t dtype;
    double* moments;
};

// T: storage type.  128-bit loads (4 fp32 / 8 half elements per thread and iteration) when the base is 16-byte aligned;
// per-thread sums of a vector in fp32, accumulated in fp64.
template <typename T>
__global__ void __launch_bounds__(256) moments_kernel(const __grid_constant__ MomentParams p) {
    constexpr int N = Vec<T>::N;
    const T* in = reinterpret_cast<const T*>(p.in);
    double s1 = 0.0, s2 = 0.0;
    const bool vec = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
    const int64_t groups = vec ? p.numel / N : 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        float v[N];
        Vec<T>::load(in + g * N, v);
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int j = 0; j < N; ++j) { a += v[j]; b += v[j] * v[j]; }
        s1 += (double)a;
        s2 += (double)b;
    }
    for (int64_t e = groups * N + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.numel; e += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)Vec<T>::get(in + e);
        s1 += v;
        s2 += v * v;
    }
    publish_sums(s1, s2, p.moments);
}

__global__ void __launch_bounds__(256) moments_kernel_f64(const __grid_constant__ MomentParams p) {
    double s1 = 0.0, s2 = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.numel; e += (int64_t)gridDim.x * blockDim.x) {
        const double v = reinterpret_cast<const double*>(p.in)[e];
        s1 += v;
        s2 += v * v;
    }
    publish_sums(s1, s2, p.moments);
}

// Error norms of an embedded Runge-Kutta pair in one pass: sums[0] += sum |low - high|^p, sums[1] += sum |high|^p
// (p = 1 or 2).  Replaces `mean(abs(low - high) ** p)` and `mean(abs(0 - high) ** p)` - six elementwise passes, two
// reductions and two host reads per adaptive step (reference: functional.py:197-214, 437-441).
struct ErrorNormParams {
    const void* low;
    const void* high;
    int64_t numel;
    int32_t dtype, power;
    double* sums;
};

template <typename T>
__global__ void __launch_bounds__(256) error_norm_kernel(const __grid_constant__ ErrorNormParams p) {
    constexpr int N = Vec<T>::N;
    const T* low = reinterpret_cast<const T*>(p.low);
    const T* high = reinterpret_cast<const T*>(p.high);
    double s1 = 0.0, s2 = 0.0;
    const bool vec = ((reinterpret_cast<uintptr_t>(low) | reinterpret_cast<uintptr_t>(high)) & 15u) == 0;
    const int64_t groups = vec ? p.numel / N : 0;
    const bool square = p.power == 2;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        float lo[N], hi[N];
        Vec<T>::load(low + g * N, lo);
        Vec<T>::load(high + g * N, hi);
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const float d = fabsf(lo[j] - hi[j]), h = fabsf(hi[j]);
            a += square ? d * d : d;
            b += square ? h * h : h;
        }
        s1 += (double)a;
        s2 += (double)b;
    }
    for (int64_t e = groups * N + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.numel; e += (int64_t)gridDim.x * blockDim.x) {
        const double lo = (double)Vec<T>::get(low + e), hi = (double)Vec<T>::get(high + e);
        const double d = fabs(lo - hi), h = fabs(hi);
        s1 += square ? d * d : d;
        s2 += square ? h * h : h;
    }
    publish_sums(s1, s2, p.sums);
}

__global__ void __launch_bounds__(256) error_norm_kernel_f64(const __grid_constant__ ErrorNormParams p) {
    double s1 = 0.0, s2 = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.numel; e += (int64_t)gridDim.x * blockDim.x) {
        const double lo = reinterpret_cast<const double*>(p.low)[e], hi = reinterpret_cast<const double*>(p.high)[e];
        const double d = fabs(lo - hi), h = fabs(hi);
        s1 += p.power == 2 ? d * d : d;
        s2 += p.power == 2 ? h * h : h;
    }
    publish_sums(s1, s2, p.sums);
}

struct ScaleParams {
    const void* in;
    void* out;
    int64_t numel;
    int32_t in_dtype, out_dtype;
    // scale = numerator [* std(num_moments)] [/ std(moments)], each std unbiased over its count
    double numerator;
    const double* num_moments;
    int64_t num_count;
    const double* moments;
    int64_t count;
    double min_std;  // leave the data unscaled when the denominator std <= min_std (reference: `if cstd > 1e-8`)
};

__device__ __forceinline__ double std_from(const double* m, int64_t n) {
    const double mean = m[0] / (double)n;
    const double var = (m[1] - (double)n * mean * mean) / (double)(n > 1 ? n - 1 : 1);
    return sqrt(var > 0.0 ? var : 0.0);
}

__device__ __forceinline__ double scale_factor(const ScaleParams& p) {
    double scale = p.numerator;
    if (p.num_moments) scale *= std_from(p.num_moments, p.num_count);
    if (p.moments) {
        const double sd = std_from(p.moments, p.count);
        scale = sd > p.min_std ? scale / sd : 1.0;
    }
    return scale;
}

// TI -> TO with 128-bit accesses on both sides: a thread moves max(N_in, N_out) consecutive elements per iteration.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) scale_kernel(const __grid_constant__ ScaleParams p) {
    constexpr int NI = Vec<TI>::N, NO = Vec<TO>::N, N = NI > NO ? NI : NO;
    const TI* in = reinterpret_cast<const TI*>(p.in);
    TO* out = reinterpret_cast<TO*>(p.out);
    const float fs = (float)scale_factor(p);
    const bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    const int64_t groups = vec ? p.numel / N : 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        float v[N];
#pragma unroll
        for (int k = 0; k < N / NI; ++k) {
            float part[NI];
            Vec<TI>::load(in + g * N + k * NI, part);
#pragma unroll
            for (int j = 0; j < NI; ++j) v[k * NI + j] = part[j] * fs;
        }
#pragma unroll
        for (int k = 0; k < N / NO; ++k) {
            float part[NO];
#pragma unroll
            for (int j = 0; j < NO; ++j) part[j] = v[k * NO + j];
            Vec<TO>::store(out + g * N + k * NO, part);
        }
    }
    for (int64_t e = groups * N + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.numel; e += (int64_t)gridDim.x * blockDim.x)
        Vec<TO>::put(out + e, Vec<TI>::get(in + e) * fs);
}

// anything involving fp64 storage: element-wise, scaled in double
__global__ void __launch_bounds__(256) scale_kernel_any(const __grid_constant__ ScaleParams p) {
    const double scale = scale_factor(p);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.numel; e += (int64_t)gridDim.x * blockDim.x) {
        const double v = p.in_dtype == SKR_F64 ? reinterpret_cast<const double*>(p.in)[e] : (double)load1(p.in, p.in_dtype, e);
        if (p.out_dtype == SKR_F64) reinterpret_cast<double*>(p.out)[e] = v * scale;
        else store1(p.out, p.out_dtype, e, (float)(v * scale));
    }
}

// ---------------------------------------------------------------------------------------------------------
// Pyramid   (reference: noise.py:146-207)

struct PyramidParams {
    void* out;
    int64_t numel;
    int32_t dtype, ndim, n_levels, mode;  // mode 0: accumulate moments only, 1: write out = value / std,
                                          // mode 2: write the unnormalised value to scratch and accumulate moments
    uint64_t seed;
    int64_t shape[SKR_MAX_DIMS];
    int32_t masked[SKR_MAX_DIMS];           // 1: axis is resized by the pyramid
    // level l: source (Philox stream or supplied fp32 buffer), resized extents along the masked axes, weight
    uint64_t stream[SKR_MAX_LEVELS];
    const float* buffer[SKR_MAX_LEVELS];    // non-null: read the level from this tensor instead of Philox
    int64_t extent[SKR_MAX_LEVELS][2];      // size of the (up to two) masked axes at this level, in axis order
    float weight[SKR_MAX_LEVELS];           // strength^l (0 for skipped levels)
    uint64_t base_stream;
    const float* base_buffer;
    float* scratch;
    double* moments;
    int32_t m_axis[2];                                  // the resized axes in order (-1: only one)
    int32_t same_size[SKR_MAX_LEVELS];                  // 1: the level has the unit shape (interpolation is the identity)
    float ratio[SKR_MAX_LEVELS][2];                     // level extent / unit extent along the resized axes (fp32, like ATen)
    const float* wide[SKR_MAX_LEVELS];                  // trailing layout: the level already interpolated along the last axis
                                                        // ([slices, level height, unit width]); null: interpolate here
    int64_t lstride[SKR_MAX_LEVELS][SKR_MAX_DIMS];      // row-major strides of each level grid
};

// F.interpolate(align_corners=False) source position along one axis (ATen area_pixel_compute_source_index)
__device__ __forceinline__ void source_index(int64_t dst, int64_t in_size, int64_t out_size, int64_t& i0, int64_t& i1, float& w1) {
    const float scale = (float)in_size / (float)out_size;
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.0f ? 0.0f : src;
    i0 = (int64_t)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    w1 = src - (float)i0;
}

__device__ __forceinline__ float level_value(const PyramidParams& p, const Philox& ph, int l, int64_t linear) {
    return p.buffer[l] ? p.buffer[l][linear] : normal_at(ph, (uint64_t)linear, p.stream[l]);
}

// One element of base + sum_l weight_l * upsample(level_l).  I is the index type: 32-bit arithmetic when the
// tensor has fewer than 2^31 elements (the integer divisions of the coordinate decomposition dominate otherwise).
template <typename I>
__device__ __forceinline__ float pyramid_value(const PyramidParams& p, const Philox& ph, int64_t e) {
    I idx[SKR_MAX_DIMS];
    I rem = (I)e;
    for (int d = p.ndim - 1; d >= 0; --d) {
        const I size = (I)p.shape[d];
        const I q = rem / size;
        idx[d] = rem - q * size;
        rem = q;
    }
    const int m0 = p.m_axis[0], m1 = p.m_axis[1];

    float pyramid = 0.0f;  // the reference starts from a zero tensor and adds the kept levels in order
    for (int l = 0; l < p.n_levels; ++l) {
        const float wl = p.weight[l];
        if (wl == 0.0f) continue;
        if (p.same_size[l]) {  // a level of the unit's own size: interpolation is the identity (weights 1 and 0)
            pyramid += level_value(p, ph, l, (int64_t)e) * wl;
            continue;
        }
        I outer = 0;  // offset of this element's slice inside the level grid (all but the resized axes)
        for (int d = 0; d < p.ndim; ++d)
            if (!p.masked[d]) outer += idx[d] * (I)p.lstride[l][d];
        float v;
        int64_t h0, h1;
        float wh;
        source_index((int64_t)idx[m0], p.extent[l][0], p.shape[m0], h0, h1, wh);
        const I s0 = (I)p.lstride[l][m0];
        if (m1 < 0) {
            v = (1.0f - wh) * level_value(p, ph, l, outer + (I)h0 * s0) + wh * level_value(p, ph, l, outer + (I)h1 * s0);
        } else {
            int64_t w0, w1;
            float ww;
            source_index((int64_t)idx[m1], p.extent[l][1], p.shape[m1], w0, w1, ww);
            const I s1 = (I)p.lstride[l][m1];
            const I r0 = outer + (I)h0 * s0, r1 = outer + (I)h1 * s0;
            const float top = (1.0f - ww) * level_value(p, ph, l, r0 + (I)w0 * s1) + ww * level_value(p, ph, l, r0 + (I)w1 * s1);
            const float bot = (1.0f - ww) * level_value(p, ph, l, r1 + (I)w0 * s1) + ww * level_value(p, ph, l, r1 + (I)w1 * s1);
            v = (1.0f - wh) * top + wh * bot;
        }
        pyramid += v * wl;
    }
    const float base = p.base_buffer ? p.base_buffer[e] : normal_at(ph, (uint64_t)e, p.base_stream);
    return base + pyramid;
}

template <typename I>
__global__ void __launch_bounds__(256) pyramid_kernel(const __grid_constant__ PyramidParams p) {
    const Philox ph(p.seed);
    double s1 = 0.0, s2 = 0.0;
    float inv_std = 1.0f;
    if (p.mode == 1) inv_std = (float)(1.0 / std_from(p.moments, p.numel));
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.numel; e += (int64_t)gridDim.x * blockDim.x) {
        const float v = pyramid_value<I>(p, ph, e);
        if (p.mode != 1) { s1 += v; s2 += (double)v * v; }
        if (p.mode == 1) store1(p.out, p.dtype, e, v * inv_std);
        else if (p.mode == 2) p.scratch[e] = v;  // may alias base_buffer: element e is read before it is written
    }
    if (p.mode != 1) publish_sums(s1, s2, p.moments);
}

// F.interpolate(align_corners=False) source position along one axis in 32-bit arithmetic (extents < 2^31)
__device__ __forceinline__ void source_index32(int dst, int in_size, float scale, int& i0, int& i1, float& w1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.0f ? 0.0f : src;
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    w1 = src - (float)i0;
}

// All the coarse level grids of one pyramid in ONE launch: blockIdx.y = level, the level's normals written as fp32
// (the values skr_noise_fill writes for (seed, stream[level])).
struct LevelsFillParams {
    uint64_t seed;
    uint64_t stream[SKR_MAX_LEVELS];
    float* out[SKR_MAX_LEVELS];
    int32_t numel[SKR_MAX_LEVELS];
};

__global__ void __launch_bounds__(256) levels_fill_kernel(const __grid_constant__ LevelsFillParams p) {
    const int l = blockIdx.y;
    const int32_t numel = p.numel[l];
    float* const out = p.out[l];
    if (!out) return;
    const Philox ph(p.seed);
    const int32_t groups = (numel + 3) >> 2;
    for (int32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
        float z[4];
        normal4(ph((uint64_t)g, p.stream[l]), z);
        const int32_t first = g << 2;
        if (first + 4 <= numel) *reinterpret_cast<float4*>(out + first) = make_float4(z[0], z[1], z[2], z[3]);  // grids are 16-byte aligned
        else for (int j = 0; j < 4 && first + j < numel; ++j) out[first + j] = z[j];
    }
}

// Bilinear interpolation is separable: a coarse level is first stretched along the last axis to the unit's width
// (this kernel: [rows, level width] -> [rows, unit width], rows = slices x level height, all levels in one launch,
// blockIdx.y = level), so the composition only interpolates between two ROWS - two 128-bit loads and four lerps per
// level and group of four elements instead of sixteen gathers and four index computations.
struct LevelsWidenParams {
    const float* in[SKR_MAX_LEVELS];
    float* out[SKR_MAX_LEVELS];
    int32_t rows[SKR_MAX_LEVELS];
    int32_t in_width[SKR_MAX_LEVELS];
    float ratio[SKR_MAX_LEVELS];
    int32_t width;  // unit width, a multiple of 4
};

__global__ void __launch_bounds__(256) levels_widen_kernel(const __grid_constant__ LevelsWidenParams p) {
    const int l = blockIdx.y;
    const float* const in = p.in[l];
    if (!in) return;
    float* const out = p.out[l];
    const int32_t width = p.width, lw = p.in_width[l];
    const float ratio = p.ratio[l];
    const int32_t groups = p.rows[l] * (width >> 2);
    for (int32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
        const int32_t e0 = g << 2;
        const int32_t row = e0 / width;
        const int32_t x = e0 - row * width;
        const float* const line = in + row * lw;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int w0, w1;
            float ww;
            source_index32(x + j, lw, ratio, w0, w1, ww);
            v[j] = (1.0f - ww) * line[w0] + ww * line[w1];
        }
        *reinterpret_cast<float4*>(out + e0) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// The composition pass, four consecutive elements of the last axis per thread (last extent % 4 == 0, fewer than 2^31
// elements).  The base draw and every level of the unit's own size (level 0 always is) are drawn in registers from
// their Philox streams - element 4g..4g+3 of a stream is block g - unless a buffer is supplied; coarse levels are
// interpolated from their grids (small: they stay in L1 / L2).  Writes the unnormalised field to `scratch` and
// accumulates its sum / sum of squares.  Same arithmetic, in the same order, as pyramid_value.
__global__ void __launch_bounds__(256) pyramid_compose4_kernel(const __grid_constant__ PyramidParams p) {
    const int last = p.ndim - 1;
    const int m0 = p.m_axis[0], m1 = p.m_axis[1];
    const bool last_resized = p.masked[last] != 0;
    const int other = last_resized ? (m1 >= 0 ? m0 : -1) : -1;  // the resized axis that is not the last one
    const int32_t groups = (int32_t)(p.numel >> 2);
    const Philox ph(p.seed);
    double s1 = 0.0, s2 = 0.0;
    for (int32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
        const int32_t e0 = g << 2;
        int32_t idx[SKR_MAX_DIMS];
        int32_t rem = e0;
        for (int d = last; d >= 0; --d) {
            const int32_t size = (int32_t)p.shape[d];
            const int32_t q = rem / size;
            idx[d] = rem - q * size;
            rem = q;
        }
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int l = 0; l < p.n_levels; ++l) {
            const float wl = p.weight[l];
            if (wl == 0.0f) continue;
            const float* grid = p.buffer[l];
            if (p.same_size[l]) {
                float q[4];
                if (grid) {
                    const float4 v = *reinterpret_cast<const float4*>(grid + e0);
                    q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
                } else {
                    normal4(ph((uint64_t)g, p.stream[l]), q);
                }
                acc[0] += q[0] * wl; acc[1] += q[1] * wl; acc[2] += q[2] * wl; acc[3] += q[3] * wl;
                continue;
            }
            int32_t outer = 0;  // slice offset from the axes that are neither resized nor the last one
            for (int d = 0; d < last; ++d)
                if (!p.masked[d]) outer += idx[d] * (int32_t)p.lstride[l][d];
            if (!last_resized) {
                // the four elements sit in four consecutive slices: same interpolation footprint in each
                int h0, h1, w0 = 0, w1 = 0;
                float wh, ww = 0.0f;
                source_index32(idx[m0], (int)p.extent[l][0], p.ratio[l][0], h0, h1, wh);
                if (m1 >= 0) source_index32(idx[m1], (int)p.extent[l][1], p.ratio[l][1], w0, w1, ww);
                const int32_t sh = (int32_t)p.lstride[l][m0], sw = m1 >= 0 ? (int32_t)p.lstride[l][m1] : 0;
                const int32_t sl = (int32_t)p.lstride[l][last];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int32_t o = outer + (idx[last] + j) * sl;
                    float v;
                    if (m1 < 0) {
                        v = (1.0f - wh) * grid[o + h0 * sh] + wh * grid[o + h1 * sh];
                    } else {
                        const int32_t r0 = o + h0 * sh, r1 = o + h1 * sh;
                        const float top = (1.0f - ww) * grid[r0 + w0 * sw] + ww * grid[r0 + w1 * sw];
                        const float bot = (1.0f - ww) * grid[r1 + w0 * sw] + ww * grid[r1 + w1 * sw];
                        v = (1.0f - wh) * top + wh * bot;
                    }
                    acc[j] += v * wl;
                }
            } else {
                int h0 = 0, h1 = 0;
                float wh = 0.0f;
                int32_t r0 = outer, r1 = outer;
                if (other >= 0) {
                    source_index32(idx[other], (int)p.extent[l][0], p.ratio[l][0], h0, h1, wh);
                    const int32_t sh = (int32_t)p.lstride[l][other];
                    r0 = outer + h0 * sh;
                    r1 = outer + h1 * sh;
                }
                const int which = other >= 0 ? 1 : 0;
                const int extent_last = (int)p.extent[l][which];
                const float ratio_last = p.ratio[l][which];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int w0, w1;
                    float ww;
                    source_index32(idx[last] + j, extent_last, ratio_last, w0, w1, ww);
                    float v;
                    if (other < 0) {
                        v = (1.0f - ww) * grid[outer + w0] + ww * grid[outer + w1];
                    } else {
                        const float top = (1.0f - ww) * grid[r0 + w0] + ww * grid[r0 + w1];
                        const float bot = (1.0f - ww) * grid[r1 + w0] + ww * grid[r1 + w1];
                        v = (1.0f - wh) * top + wh * bot;
                    }
                    acc[j] += v * wl;
                }
            }
        }
        float b[4];
        if (p.base_buffer) {
            const float4 q = *reinterpret_cast<const float4*>(p.base_buffer + e0);
            b[0] = q.x; b[1] = q.y; b[2] = q.z; b[3] = q.w;
        } else {
            normal4(ph((uint64_t)g, p.base_stream), b);
        }
        const float4 v = make_float4(b[0] + acc[0], b[1] + acc[1], b[2] + acc[2], b[3] + acc[3]);
        *reinterpret_cast<float4*>(p.scratch + e0) = v;
        s1 += (double)((v.x + v.y) + (v.z + v.w));
        s2 += (double)((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
    }
    publish_sums(s1, s2, p.moments);
}

// The same composition for the usual layout - the resized axes are the trailing ones, everything before them only
// selects a slice: the tensor is [slices, H, W] (AXES == 2) or [slices, W] (AXES == 1).  Coordinates live in
// registers (no index arrays), two or one integer divisions per four elements.
template <int AXES>
__global__ void __launch_bounds__(256) pyramid_compose_trailing_kernel(const __grid_constant__ PyramidParams p) {
    const int32_t groups = (int32_t)(p.numel >> 2);
    const int32_t width = (int32_t)p.shape[AXES], height = AXES == 2 ? (int32_t)p.shape[1] : 1;
    const Philox ph(p.seed);
    double s1 = 0.0, s2 = 0.0;
    for (int32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
        const int32_t e0 = g << 2;
        const int32_t row = e0 / width;          // slice * height + y
        const int32_t x = e0 - row * width;
        int32_t slice = row, y = 0;
        if constexpr (AXES == 2) {
            slice = row / height;
            y = row - slice * height;
        }
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int l = 0; l < p.n_levels; ++l) {
            const float wl = p.weight[l];
            if (wl == 0.0f) continue;
            const float* grid = p.buffer[l];
            if (p.same_size[l]) {
                float q[4];
                if (grid) {
                    const float4 v = *reinterpret_cast<const float4*>(grid + e0);
                    q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
                } else {
                    normal4(ph((uint64_t)g, p.stream[l]), q);
                }
                acc[0] += q[0] * wl; acc[1] += q[1] * wl; acc[2] += q[2] * wl; acc[3] += q[3] * wl;
                continue;
            }
            if (const float* wide = p.wide[l]) {
                // already at the unit's width: interpolate between two rows (AXES == 2) or just add (AXES == 1)
                float4 top;
                if constexpr (AXES == 2) {
                    int h0, h1;
                    float wh;
                    source_index32(y, (int)p.extent[l][0], p.ratio[l][0], h0, h1, wh);
                    const int32_t base = slice * (int32_t)p.extent[l][0];
                    top = *reinterpret_cast<const float4*>(wide + (base + h0) * width + x);
                    const float4 bot = *reinterpret_cast<const float4*>(wide + (base + h1) * width + x);
                    top.x = (1.0f - wh) * top.x + wh * bot.x;
                    top.y = (1.0f - wh) * top.y + wh * bot.y;
                    top.z = (1.0f - wh) * top.z + wh * bot.z;
                    top.w = (1.0f - wh) * top.w + wh * bot.w;
                } else {
                    top = *reinterpret_cast<const float4*>(wide + e0);
                }
                acc[0] += top.x * wl; acc[1] += top.y * wl; acc[2] += top.z * wl; acc[3] += top.w * wl;
                continue;
            }
            const int lw = (int)p.extent[l][AXES - 1];
            const float rw = p.ratio[l][AXES - 1];
            const float* r0 = grid + slice * (int32_t)p.lstride[l][0];
            const float* r1 = r0;
            float wh = 0.0f;
            if constexpr (AXES == 2) {
                int h0, h1;
                source_index32(y, (int)p.extent[l][0], p.ratio[l][0], h0, h1, wh);
                r1 = r0 + h1 * lw;
                r0 = r0 + h0 * lw;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int w0, w1;
                float ww;
                source_index32(x + j, lw, rw, w0, w1, ww);
                float v = (1.0f - ww) * r0[w0] + ww * r0[w1];
                if constexpr (AXES == 2) {
                    const float bot = (1.0f - ww) * r1[w0] + ww * r1[w1];
                    v = (1.0f - wh) * v + wh * bot;
                }
                acc[j] += v * wl;
            }
        }
        float b[4];
        if (p.base_buffer) {
            const float4 q = *reinterpret_cast<const float4*>(p.base_buffer + e0);
            b[0] = q.x; b[1] = q.y; b[2] = q.z; b[3] = q.w;
        } else {
            normal4(ph((uint64_t)g, p.base_stream), b);
        }
        const float4 v = make_float4(b[0] + acc[0], b[1] + acc[1], b[2] + acc[2], b[3] + acc[3]);
        *reinterpret_cast<float4*>(p.scratch + e0) = v;
        s1 += (double)((v.x + v.y) + (v.z + v.w));
        s2 += (double)((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
    }
    publish_sums(s1, s2, p.moments);
}

// Composition AND normalisation in one launch, for units that fit the shared memory of the whole GPU (a 16x21x90x160
// video latent is 19.4 MB against 148 x 227 KB): a cooperative grid of one 1024-thread CTA per SM, every CTA owning a
// contiguous run of the unit.  Phase 1 is the composition of pyramid_compose_trailing_kernel (same arithmetic, same
// order) with the unnormalised values HELD IN SHARED MEMORY instead of written to a scratch tensor; the grid-wide sums
// are published in a fixed order; every CTA waits for the totals (generation counter, acquire load); phase 2 scales
// what it holds and stores it once, in the output's storage type.  Against composition + scale pass this removes a
// launch and a 4-byte-per-element round trip.  K float4 groups per thread and iteration (the library launches K = 2: a
// row holds an even number of groups and the row arithmetic of a level is shared by eight elements); coordinates
// advance by carries, three integer divisions per thread in total.
template <int AXES, int K, typename TO>
__global__ void __launch_bounds__(1024, 1) pyramid_resident_kernel(const __grid_constant__ PyramidParams p, const int32_t units_per_cta) {
    extern __shared__ __align__(16) float4 held[];
    __shared__ unsigned int generation;
    const int32_t width = (int32_t)p.shape[AXES], height = AXES == 2 ? (int32_t)p.shape[1] : 1;
    const int32_t upr = width / (4 * K);  // units (K groups of four elements) per row
    const int32_t units = (int32_t)(p.numel >> 2) / K;
    const int32_t first = (int32_t)blockIdx.x * units_per_cta;
    const int32_t count = units - first < units_per_cta ? (units - first > 0 ? units - first : 0) : units_per_cta;
    unsigned int* const gen = reinterpret_cast<unsigned int*>(p.moments + 3);
    if (threadIdx.x == 0) generation = load_acquire(gen);  // before this CTA's arrival: the same value in every CTA
    const Philox ph(p.seed);
    const int32_t stride = (int32_t)blockDim.x;
    int32_t xu, slice, y = 0;
    {
        const int32_t u = first + (int32_t)threadIdx.x;
        const int32_t row = u / upr;
        xu = u - row * upr;
        slice = row;
        if constexpr (AXES == 2) {
            slice = row / height;
            y = row - slice * height;
        }
    }
    const int32_t dq = stride / upr, dr = stride - dq * upr;
    const int32_t dqs = AXES == 2 ? dq / height : dq, dqy = AXES == 2 ? dq - dqs * height : 0;
    double s1 = 0.0, s2 = 0.0;
    for (int32_t i = (int32_t)threadIdx.x; i < count; i += stride) {
        const int32_t g0 = (first + i) * K;
        const int32_t x = xu * (4 * K);
        float acc[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.0f;
        for (int l = 0; l < p.n_levels; ++l) {
            const float wl = p.weight[l];
            if (wl == 0.0f) continue;
            if (p.same_size[l]) {
                const float* grid = p.buffer[l];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float q[4];
                    if (grid) {
                        const float4 v = *reinterpret_cast<const float4*>(grid + ((g0 + k) << 2));
                        q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
                    } else {
                        normal4(ph((uint64_t)(g0 + k), p.stream[l]), q);
                    }
                    acc[k][0] += q[0] * wl; acc[k][1] += q[1] * wl; acc[k][2] += q[2] * wl; acc[k][3] += q[3] * wl;
                }
                continue;
            }
            const float* wide = p.wide[l];
            if constexpr (AXES == 2) {
                int h0, h1;
                float wh;
                source_index32(y, (int)p.extent[l][0], p.ratio[l][0], h0, h1, wh);
                const int32_t base = slice * (int32_t)p.extent[l][0];
                const float* r0 = wide + (base + h0) * width + x;
                const float* r1 = wide + (base + h1) * width + x;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float4 top = *reinterpret_cast<const float4*>(r0 + 4 * k);
                    const float4 bot = *reinterpret_cast<const float4*>(r1 + 4 * k);
                    top.x = (1.0f - wh) * top.x + wh * bot.x;
                    top.y = (1.0f - wh) * top.y + wh * bot.y;
                    top.z = (1.0f - wh) * top.z + wh * bot.z;
                    top.w = (1.0f - wh) * top.w + wh * bot.w;
                    acc[k][0] += top.x * wl; acc[k][1] += top.y * wl; acc[k][2] += top.z * wl; acc[k][3] += top.w * wl;
                }
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float4 top = *reinterpret_cast<const float4*>(wide + ((g0 + k) << 2));
                    acc[k][0] += top.x * wl; acc[k][1] += top.y * wl; acc[k][2] += top.z * wl; acc[k][3] += top.w * wl;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float b[4];
            if (p.base_buffer) {
                const float4 q = *reinterpret_cast<const float4*>(p.base_buffer + ((g0 + k) << 2));
                b[0] = q.x; b[1] = q.y; b[2] = q.z; b[3] = q.w;
            } else {
                normal4(ph((uint64_t)(g0 + k), p.base_stream), b);
            }
            const float4 v = make_float4(b[0] + acc[k][0], b[1] + acc[k][1], b[2] + acc[k][2], b[3] + acc[k][3]);
            held[i * K + k] = v;
            s1 += (double)((v.x + v.y) + (v.z + v.w));
            s2 += (double)((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
        }
        // the next unit of this thread: `stride` units further on
        xu += dr;
        const int32_t carry = xu >= upr ? 1 : 0;
        xu -= carry ? upr : 0;
        if constexpr (AXES == 2) {
            y += dqy + carry;
            slice += dqs;
            if (y >= height) {
                y -= height;
                ++slice;
            }
        } else {
            slice += dq + carry;
        }
    }
    publish_sums<true>(s1, s2, p.moments);
    if (threadIdx.x == 0) {
        const unsigned int before = generation;
        while (load_acquire(gen) == before) __nanosleep(40);
    }
    __syncthreads();
    double m[2] = {__ldcg(p.moments), __ldcg(p.moments + 1)};
    const double sd = std_from(m, p.numel);
    const float fs = (float)(sd > 0.0 ? 1.0 / sd : 1.0);
    TO* const out = reinterpret_cast<TO*>(p.out) + ((int64_t)first * K << 2);
    for (int32_t i = (int32_t)threadIdx.x; i < count * K; i += stride) {
        const float4 v = held[i];
        const float scaled[4] = {v.x * fs, v.y * fs, v.z * fs, v.w * fs};
        put4<TO>(out + (i << 2), scaled);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Colored: spectral shaping of an rfftn half spectrum, in place   (reference: noise.py:285-335,379-394)

struct ShapeParams {
    float2* spectrum;     // complex64, shape = dims[0..ndim-2] x (dims[ndim-1]/2 + 1)
    double2* spectrum64;  // complex128 alternative (exactly one of the two is non-null)
    int64_t bins;         // number of complex elements
    int32_t ndim;
    int64_t dims[SKR_MAX_DIMS];  // real-space extents of the transformed axes
    float exponent_half_neg;     // -exponent / 2
    float eps_clip;
    float r_max;
};

// I: index type (32-bit below 2^31 bins: the per-axis divisions dominate the kernel otherwise)
template <typename I>
__global__ void __launch_bounds__(256) colored_shape_kernel(const __grid_constant__ ShapeParams p) {
    const int last = p.ndim - 1;
    const I last_bins = (I)(p.dims[last] / 2 + 1);
    const I bins = (I)p.bins;
    for (I b = (I)(blockIdx.x * blockDim.x + threadIdx.x); b < bins; b += (I)(gridDim.x * blockDim.x)) {
        I rem = b;
        float r2 = 0.0f;
        {
            const I q = rem / last_bins;
            const I k = rem - q * last_bins;
            rem = q;
            const float f = (float)k / (float)p.dims[last];
            r2 = f * f;
        }
        for (int d = last - 1; d >= 0; --d) {
            const I n = (I)p.dims[d];
            const I q = rem / n;
            const I k = rem - q * n;
            rem = q;
            // |fftfreq(n)|: k/n for k < ceil(n/2), else (n-k)/n
            const I kk = k < (n + 1) / 2 ? k : n - k;
            const float f = (float)kk / (float)n;
            r2 += f * f;
        }
        float r = sqrtf(r2);
        if (p.r_max > 0.0f) r = r / p.r_max;
        r = r < p.eps_clip ? p.eps_clip : r;
        const float w = powf(r, p.exponent_half_neg);
        if (p.spectrum) {
            float2 v = p.spectrum[b];
            v.x *= w;
            v.y *= w;
            p.spectrum[b] = v;
        } else {
            double2 v = p.spectrum64[b];
            v.x *= (double)w;
            v.y *= (double)w;
            p.spectrum64[b] = v;
        }
    }
}

// The same for complex64 spectra with many rows: one warp per row of the last (half) axis.  The row's index is
// decomposed once per warp instead of once per bin, the lanes walk the row with coalesced 8-byte accesses, and the
// power is exp2(e * log2(r)) on the special-function unit (relative error ~1e-6 against powf; the field is
// renormalised by its measured std afterwards).
__global__ void __launch_bounds__(256) colored_shape_rows_kernel(const __grid_constant__ ShapeParams p, const int rows, const float inv_r_max) {
    const int last = p.ndim - 1;
    const int last_bins = (int)(p.dims[last] / 2 + 1);
    const float inv_last = 1.0f / (float)p.dims[last];
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
        int rem = row;
        float row_r2 = 0.0f;
        for (int d = last - 1; d >= 0; --d) {
            const int n = (int)p.dims[d];
            const int q = rem / n;
            const int k = rem - q * n;
            rem = q;
            const int kk = k < (n + 1) / 2 ? k : n - k;
            const float f = (float)kk / (float)n;
            row_r2 += f * f;
        }
        float2* const line = p.spectrum + (int64_t)row * last_bins;
        for (int k = lane; k < last_bins; k += 32) {
            const float f = (float)k * inv_last;
            float r = sqrtf(row_r2 + f * f);
            if (p.r_max > 0.0f) r = r * inv_r_max;
            r = r < p.eps_clip ? p.eps_clip : r;
            const float w = exp2f(p.exponent_half_neg * __log2f(r));
            float2 v = line[k];
            v.x *= w;
            v.y *= w;
            line[k] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Brownian interval   (reference: noise.py:210-252 - torchsde.BrownianInterval over normalised time 0..1)
//
// A stateless Levy construction: the increment of a dyadic interval splits into its two halves as
// d/2 +- sqrt(h)/2 * Z(node), Z(node) = the Philox normals of stream (1 << 63 | heap index of the node).  A query
// W(t1) - W(t0) walks the tree from the root: while both times fall in the same half only the interval's
// increment is tracked; below the node that separates them each time descends as a Brownian bridge whose end
// values are kept RELATIVE to the separating midpoint (so fp32 keeps the precision of the small increment),
// and inside its leaf (width 2^-depth) each time is placed by one exact bridge draw.  Every (seed, t0, t1)
// therefore gives the same tensor whenever and wherever it is asked for, increments over adjoining intervals
// add up, and increments over disjoint intervals are independent.  The host builds the walk; the kernel
// evaluates it for four elements per thread, one Philox block per visited node.

struct BrownianNode {
    uint64_t stream;
    float scale;    // sqrt(width)/2 of the node's interval (signed in the shared prefix: + left half, - right half)
    int32_t right;  // the walk continues in the right half
};

struct BrownianWalk {
    BrownianNode node[SKR_BROWNIAN_MAX_DEPTH];
    uint64_t leaf_stream;
    float frac, sd;  // position inside the leaf and the bridge deviation there
    int32_t n;
    int32_t pad;
};

struct BrownianParams {
    void* out;
    int64_t numel;  // elements per batch item; item i (blockIdx.y) has its own seed and starts at out + i * numel
    uint64_t seed[SKR_MAX_PHILOX_ITEMS];
    uint64_t root_stream;
    BrownianNode prefix[SKR_BROWNIAN_MAX_DEPTH];
    BrownianNode split;
    BrownianWalk from, to;
    int32_t n_prefix, same_leaf, dtype, pad;
    float out_scale;
};

__device__ __forceinline__ void brownian_descend(const Philox& ph, uint64_t g, const BrownianWalk& w, float (&lo)[4], float (&hi)[4],
                                                 float (&at)[4]) {
    float z[4];
    for (int k = 0; k < w.n; ++k) {
        normal4(ph(g, w.node[k].stream), z);
        const float s = w.node[k].scale;
        const bool right = w.node[k].right != 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float mid = 0.5f * (lo[j] + hi[j]) + s * z[j];
            lo[j] = right ? mid : lo[j];
            hi[j] = right ? hi[j] : mid;
        }
    }
    normal4(ph(g, w.leaf_stream), z);
#pragma unroll
    for (int j = 0; j < 4; ++j) at[j] = (lo[j] + w.frac * (hi[j] - lo[j])) + w.sd * z[j];
}

__global__ void __launch_bounds__(256) brownian_kernel(const __grid_constant__ BrownianParams p) {
    const Philox ph(p.seed[blockIdx.y]);
    const int esize = p.dtype == SKR_F32 ? 4 : p.dtype == SKR_F64 ? 8 : 2;
    void* const out = reinterpret_cast<char*>(p.out) + (int64_t)blockIdx.y * p.numel * esize;
    const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    const int64_t groups = (p.numel + 3) >> 2;
    for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t g = (uint64_t)gi;
        float d[4], z[4], v[4];
        normal4(ph(g, p.root_stream), d);  // W(1) - W(0)
        for (int k = 0; k < p.n_prefix; ++k) {
            normal4(ph(g, p.prefix[k].stream), z);
            const float s = p.prefix[k].scale;
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = 0.5f * d[j] + s * z[j];
        }
        if (p.same_leaf) {
            // both times inside one leaf: frac / sd hold the differences of the two placements
            normal4(ph(g, p.to.leaf_stream), z);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (p.to.frac * d[j] + p.to.sd * z[j]) * p.out_scale;
        } else {
            normal4(ph(g, p.split.stream), z);
            float lo[4], hi[4], w_from[4], w_to[4];
            // left half [l, m]: values relative to W(m)
#pragma unroll
            for (int j = 0; j < 4; ++j) { lo[j] = -(0.5f * d[j] + p.split.scale * z[j]); hi[j] = 0.0f; }
            float right_end[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) right_end[j] = 0.5f * d[j] - p.split.scale * z[j];
            brownian_descend(ph, g, p.from, lo, hi, w_from);
#pragma unroll
            for (int j = 0; j < 4; ++j) { lo[j] = 0.0f; hi[j] = right_end[j]; }
            brownian_descend(ph, g, p.to, lo, hi, w_to);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (w_to[j] - w_from[j]) * p.out_scale;
        }
        const int64_t first = gi << 2;
        if (aligned && first + 4 <= p.numel) {
            store4(out, p.dtype, first, v);
        } else {
            for (int j = 0; j < 4 && first + j < p.numel; ++j) store1(out, p.dtype, first + j, v[j]);
        }
    }
}

static unsigned grid_for(int64_t work_items, int threads) {
    const int64_t sms = sm_count_or(148);
    int64_t blocks = (work_items + threads - 1) / threads;
    int64_t cap = sms * 8;
    if (cap > SKR_MOMENT_BLOCKS) cap = SKR_MOMENT_BLOCKS;  // kernels that publish grid-wide sums keep one partial per block
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

static int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "%s launch: %s", what, cudaGetErrorString(e));
    count_launch(2);
    return 0;
}


// The resident composition (pyramid_resident_kernel): returns 0 when launched, a negative value when this unit / device
// does not qualify (the caller then runs composition + scale pass), a positive CUDA status when the launch failed.
template <int AXES, int K, typename TO>
static int launch_resident_as(PyramidParams& p, int sms, int max_smem, cudaStream_t s) {
    const int64_t units = (p.numel >> 2) / K;
    const int64_t per_cta = (units + sms - 1) / sms;
    const int64_t bytes = per_cta * K * (int64_t)sizeof(float4);
    if (bytes + 2048 > max_smem) return -1;  // static shared memory of the kernel: 1.5 KB
    auto kernel = pyramid_resident_kernel<AXES, K, TO>;
    static std::atomic<uint64_t> prepared{0};  // one bit per device: the shared-memory opt-in has been set
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (!(prepared.load(std::memory_order_acquire) >> dev & 1u)) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 2048) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
        prepared.fetch_or((uint64_t)1 << dev, std::memory_order_release);
    }
    int32_t units_per_cta = (int32_t)per_cta;
    void* args[] = {&p, &units_per_cta};
    // cooperative: the grid-wide wait inside the kernel needs every CTA resident, and this launch guarantees it
    const cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3((unsigned)sms), dim3(1024), args, (size_t)bytes, s);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported ? -1 : fail((int)e, "pyramid resident launch: %s", cudaGetErrorString(e));
    }
    count_launch(2);
    return 0;
}

template <int AXES, int K>
static int launch_resident_typed(PyramidParams& p, int sms, int max_smem, cudaStream_t s) {
    switch (p.dtype) {
        case SKR_F32: return launch_resident_as<AXES, K, float>(p, sms, max_smem, s);
        case SKR_BF16: return launch_resident_as<AXES, K, __nv_bfloat16>(p, sms, max_smem, s);
        case SKR_F16: return launch_resident_as<AXES, K, __half>(p, sms, max_smem, s);
        default: return -1;
    }
}

// p is the collapsed trailing-layout descriptor ([slices, H, W] or [slices, W]) the composition kernel would get.
static int launch_resident(PyramidParams& p, int axes, cudaStream_t s) {
    if (env_int("SKR_NO_RESIDENT_PYRAMID", 0)) return -1;
    int dev = 0, coop = 0, sms = 0, max_smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    static std::atomic<int> cached[64][3];  // cooperative launch, SM count, opt-in shared memory (+1: 0 = not read yet)
    if (dev < 0 || dev >= 64) return -1;
    if (cached[dev][1].load(std::memory_order_acquire) == 0) {
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cached[dev][0].store(coop, std::memory_order_relaxed);
        cached[dev][2].store(max_smem, std::memory_order_relaxed);
        cached[dev][1].store(sms > 0 ? sms : -1, std::memory_order_release);
    }
    coop = cached[dev][0].load(std::memory_order_relaxed);
    sms = cached[dev][1].load(std::memory_order_relaxed);
    max_smem = cached[dev][2].load(std::memory_order_relaxed);
    if (!coop || sms < 1 || sms > SKR_MOMENT_BLOCKS || max_smem < 64 * 1024) return -1;
    for (int l = 0; l < p.n_levels; ++l)
        if (p.weight[l] != 0.0f && !p.same_size[l] && !p.wide[l]) return -1;  // a coarse level that was not widened
    const int width = (int)p.shape[axes];
    const int align = p.dtype == SKR_F32 ? 15 : 7;
    if ((reinterpret_cast<uintptr_t>(p.out) & align) != 0) return -1;
    // two groups per thread and iteration: a row must hold an even number of groups (width % 8 == 0; every latent
    // shape in use does).  Other widths keep composition + scale pass.
    if ((width & 7) != 0) return -1;
    return axes == 2 ? launch_resident_typed<2, 2>(p, sms, max_smem, s) : launch_resident_typed<1, 2>(p, sms, max_smem, s);
}
}  // namespace skr

extern "C" {

int skr_noise_fill(void* out, int32_t dtype, int64_t numel, uint64_t seed, uint64_t stream, const skr_offset* offset, double* moments,
                   void* cuda_stream) {
    using namespace skr;
    if (numel < 0) return fail(SKR_E_RANGE, "negative numel");
    if (dtype < 0 || dtype > SKR_F16) return fail(SKR_E_DTYPE, "unknown dtype %d", dtype);
    if (numel == 0) return 0;
    if (!out) return fail(SKR_E_NULL, "null output");
    FillParams p;
    memset(&p, 0, sizeof(p));
    p.out = out; p.numel = numel; p.seed = seed; p.stream = stream; p.dtype = dtype; p.moments = moments;
    p.aligned = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    if (offset && offset->scale != 0.0) {
        if (offset->ndim < 1 || offset->ndim > SKR_MAX_DIMS) return fail(SKR_E_SHAPE, "offset ndim %d out of range", offset->ndim);
        int64_t total = 1;
        for (int d = 0; d < offset->ndim; ++d) {
            if (offset->shape[d] < 1) return fail(SKR_E_SHAPE, "offset shape[%d] < 1", d);
            p.shape[d] = offset->shape[d];
            p.keep[d] = offset->keep[d] ? 1 : 0;
            total *= offset->shape[d];
        }
        if (total != numel) return fail(SKR_E_SHAPE, "offset shape does not multiply to numel");
        p.ndim = offset->ndim;
        p.offset_scale = (float)offset->scale;
        p.offset_stream = offset->stream;
        // kept axes first, broadcast axes after: the offset index is the element index divided by the broadcast extent
        int d = 0;
        while (d < offset->ndim && p.keep[d]) ++d;
        int64_t inner = 1;
        bool prefix = true;
        for (int k = d; k < offset->ndim; ++k) {
            prefix = prefix && !p.keep[k];
            inner *= p.shape[k];
        }
        p.inner = prefix ? inner : 0;
    }
    const unsigned grid = grid_for((numel + 3) / 4, 256);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    const bool shifted = p.offset_scale != 0.0f;
    switch (dtype) {
        case SKR_F32: launch_fill<float>(p, shifted, moments != nullptr, grid, s); break;
        case SKR_BF16: launch_fill<__nv_bfloat16>(p, shifted, moments != nullptr, grid, s); break;
        case SKR_F16: launch_fill<__half>(p, shifted, moments != nullptr, grid, s); break;
        default: launch_fill<double>(p, shifted, moments != nullptr, grid, s); break;
    }
    return check_launch("noise fill");
}

int skr_noise_fill_batch(void* out, int32_t dtype, const skr_philox* keys, void* cuda_stream) {
    using namespace skr;
    if (!keys) return fail(SKR_E_NULL, "null keys");
    if (dtype < 0 || dtype > SKR_F16) return fail(SKR_E_DTYPE, "unknown dtype %d", dtype);
    if (keys->n_items < 1 || keys->n_items > SKR_MAX_PHILOX_ITEMS) return fail(SKR_E_RANGE, "n_items %d out of range", keys->n_items);
    if (keys->item_numel < 0) return fail(SKR_E_RANGE, "negative item_numel");
    if (keys->offset_scale != 0.0f && (keys->offset_inner < 1 || keys->item_numel % keys->offset_inner != 0))
        return fail(SKR_E_SHAPE, "offset_inner %lld does not divide item_numel", (long long)keys->offset_inner);
    const int64_t numel = keys->item_numel * keys->n_items;
    if (numel == 0) return 0;
    if (!out) return fail(SKR_E_NULL, "null output");
    BatchFillParams p;
    memset(&p, 0, sizeof(p));
    p.out = out; p.numel = numel; p.dtype = dtype;
    p.aligned = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    fill_kphilox(&p.keys, keys, 1);
    const unsigned grid = grid_for((numel + 3) / 4, 256);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    switch (dtype) {
        case SKR_F32: batch_fill_kernel<float><<<grid, 256, 0, s>>>(p); break;
        case SKR_BF16: batch_fill_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p); break;
        case SKR_F16: batch_fill_kernel<__half><<<grid, 256, 0, s>>>(p); break;
        default: batch_fill_kernel<double><<<grid, 256, 0, s>>>(p); break;
    }
    return check_launch("noise batch fill");
}

int skr_noise_brownian_batch(void* out, int32_t dtype, const uint64_t* seeds, int32_t n_items, int64_t item_numel, double t0, double t1,
                             int32_t depth, double out_scale, void* cuda_stream) {
    using namespace skr;
    if (item_numel < 0) return fail(SKR_E_RANGE, "negative numel");
    if (dtype < 0 || dtype > SKR_F16) return fail(SKR_E_DTYPE, "unknown dtype %d", dtype);
    if (n_items < 1 || n_items > SKR_MAX_PHILOX_ITEMS) return fail(SKR_E_RANGE, "n_items %d out of range", n_items);
    if (!seeds) return fail(SKR_E_NULL, "null seeds");
    if (depth < 1 || depth > SKR_BROWNIAN_MAX_DEPTH) return fail(SKR_E_RANGE, "brownian depth %d outside 1..%d", depth, SKR_BROWNIAN_MAX_DEPTH);
    if (!(t0 >= 0.0 && t1 <= 1.0 && t0 < t1)) return fail(SKR_E_RANGE, "brownian interval needs 0 <= t0 < t1 <= 1");
    if (item_numel == 0) return 0;
    if (!out) return fail(SKR_E_NULL, "null output");
    const int64_t numel = item_numel;
    BrownianParams p;
    memset(&p, 0, sizeof(p));
    p.out = out; p.numel = numel; p.dtype = dtype; p.out_scale = (float)out_scale;
    for (int i = 0; i < n_items; ++i) p.seed[i] = seeds[i];
    const uint64_t tree = 1ull << 63, leaf = 1ull << 62;
    p.root_stream = tree;  // heap index 0 is unused by the nodes (the root interval is index 1)
    // interval ends are k / 2^level: exact in double for every depth allowed here
    double l = 0.0, r = 1.0;
    uint64_t node = 1;
    int level = 0;
    for (; level < depth; ++level) {
        const double m = 0.5 * (l + r);
        const bool from_right = t0 >= m, to_right = t1 >= m;
        if (from_right != to_right) break;
        const float s = (float)(0.5 * sqrt(r - l));
        p.prefix[p.n_prefix++] = BrownianNode{tree | node, from_right ? -s : s, from_right ? 1 : 0};
        if (from_right) { l = m; node = 2 * node + 1; } else { r = m; node = 2 * node; }
    }
    if (level == depth) {
        const double h = r - l;
        p.same_leaf = 1;
        p.to.leaf_stream = tree | leaf | node;
        p.to.frac = (float)((t1 - t0) / h);
        p.to.sd = (float)(sqrt((t1 - l) * (r - t1) / h) - sqrt((t0 - l) * (r - t0) / h));
    } else {
        const double m = 0.5 * (l + r);
        p.split = BrownianNode{tree | node, (float)(0.5 * sqrt(r - l)), 0};
        struct Side { BrownianWalk* walk; double t, l, r; uint64_t node; } sides[2] = {{&p.from, t0, l, m, 2 * node}, {&p.to, t1, m, r, 2 * node + 1}};
        for (Side& side : sides) {
            for (int k = level + 1; k < depth; ++k) {
                const double mid = 0.5 * (side.l + side.r);
                const bool right = side.t >= mid;
                side.walk->node[side.walk->n++] = BrownianNode{tree | side.node, (float)(0.5 * sqrt(side.r - side.l)), right ? 1 : 0};
                if (right) { side.l = mid; side.node = 2 * side.node + 1; } else { side.r = mid; side.node = 2 * side.node; }
            }
            const double h = side.r - side.l;
            side.walk->leaf_stream = tree | leaf | side.node;
            side.walk->frac = (float)((side.t - side.l) / h);
            side.walk->sd = (float)sqrt((side.t - side.l) * (side.r - side.t) / h);
        }
    }
    unsigned blocks = grid_for((numel + 3) / 4, 256);
    if (n_items > 1) blocks = (blocks + n_items - 1) / n_items;  // the grid cap is for the whole launch
    brownian_kernel<<<dim3(blocks, (unsigned)n_items), 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(p);
    return check_launch("brownian interval");
}

int skr_noise_brownian(void* out, int32_t dtype, int64_t numel, uint64_t seed, double t0, double t1, int32_t depth, double out_scale,
                       void* cuda_stream) {
    return skr_noise_brownian_batch(out, dtype, &seed, 1, numel, t0, t1, depth, out_scale, cuda_stream);
}

int skr_noise_moments(const void* in, int32_t dtype, int64_t numel, double* moments, void* cuda_stream) {
    using namespace skr;
    if (numel < 0) return fail(SKR_E_RANGE, "negative numel");
    if (dtype < 0 || dtype > SKR_F16) return fail(SKR_E_DTYPE, "unknown dtype %d", dtype);
    if (!moments) return fail(SKR_E_NULL, "null moments");
    if (numel == 0) return 0;
    if (!in) return fail(SKR_E_NULL, "null input");
    MomentParams p{in, numel, dtype, moments};
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    const unsigned grid = grid_for(numel, 256 * 8);
    switch (dtype) {
        case SKR_F32: moments_kernel<float><<<grid, 256, 0, s>>>(p); break;
        case SKR_BF16: moments_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p); break;
        case SKR_F16: moments_kernel<__half><<<grid, 256, 0, s>>>(p); break;
        default: moments_kernel_f64<<<grid, 256, 0, s>>>(p); break;
    }
    return check_launch("noise moments");
}

int skr_error_norms(const void* low, const void* high, int32_t dtype, int64_t numel, int32_t power, double* sums, void* cuda_stream) {
    using namespace skr;
    if (numel < 0) return fail(SKR_E_RANGE, "negative numel");
    if (dtype < 0 || dtype > SKR_F16) return fail(SKR_E_DTYPE, "unknown dtype %d", dtype);
    if (power != 1 && power != 2) return fail(SKR_E_RANGE, "power must be 1 (mean absolute) or 2 (mean squared), got %d", power);
    if (!sums) return fail(SKR_E_NULL, "null sums");
    if (numel == 0) return 0;
    if (!low || !high) return fail(SKR_E_NULL, "null input");
    ErrorNormParams p{low, high, numel, dtype, power, sums};
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    const unsigned grid = grid_for(numel, 256 * 8);
    switch (dtype) {
        case SKR_F32: error_norm_kernel<float><<<grid, 256, 0, s>>>(p); break;
        case SKR_BF16: error_norm_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p); break;
        case SKR_F16: error_norm_kernel<__half><<<grid, 256, 0, s>>>(p); break;
        default: error_norm_kernel_f64<<<grid, 256, 0, s>>>(p); break;
    }
    return check_launch("error norms");
}

int skr_noise_scale(const void* in, int32_t in_dtype, void* out, int32_t out_dtype, int64_t numel, double numerator,
                    const double* num_moments, int64_t num_count, const double* moments, int64_t count, double min_std, void* cuda_stream) {
    using namespace skr;
    if (numel < 0) return fail(SKR_E_RANGE, "negative numel");
    if (in_dtype < 0 || in_dtype > SKR_F16 || out_dtype < 0 || out_dtype > SKR_F16) return fail(SKR_E_DTYPE, "unknown dtype");
    if (numel == 0) return 0;
    if (!in || !out) return fail(SKR_E_NULL, "null tensor");
    ScaleParams p{in, out, numel, in_dtype, out_dtype, numerator, num_moments, num_count, moments, count, min_std};
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    const unsigned grid = grid_for(numel, 256 * 8);
#define SKR_SCALE_CASE(TI, TO) scale_kernel<TI, TO><<<grid, 256, 0, s>>>(p)
    if (in_dtype == SKR_F64 || out_dtype == SKR_F64) scale_kernel_any<<<grid, 256, 0, s>>>(p);
    else if (in_dtype == SKR_F32 && out_dtype == SKR_F32) SKR_SCALE_CASE(float, float);
    else if (in_dtype == SKR_F32 && out_dtype == SKR_BF16) SKR_SCALE_CASE(float, __nv_bfloat16);
    else if (in_dtype == SKR_F32 && out_dtype == SKR_F16) SKR_SCALE_CASE(float, __half);
    else if (in_dtype == SKR_BF16 && out_dtype == SKR_BF16) SKR_SCALE_CASE(__nv_bfloat16, __nv_bfloat16);
    else if (in_dtype == SKR_BF16 && out_dtype == SKR_F32) SKR_SCALE_CASE(__nv_bfloat16, float);
    else if (in_dtype == SKR_F16 && out_dtype == SKR_F16) SKR_SCALE_CASE(__half, __half);
    else if (in_dtype == SKR_F16 && out_dtype == SKR_F32) SKR_SCALE_CASE(__half, float);
    else if (in_dtype == SKR_BF16 && out_dtype == SKR_F16) SKR_SCALE_CASE(__nv_bfloat16, __half);
    else SKR_SCALE_CASE(__half, __nv_bfloat16);
#undef SKR_SCALE_CASE
    return check_launch("noise scale");
}

int skr_noise_pyramid(void* out, int32_t dtype, const skr_pyramid* desc, double* moments, void* cuda_stream) {
    using namespace skr;
    if (!desc || !moments) return fail(SKR_E_NULL, "null descriptor / moments");
    if (dtype < 0 || dtype > SKR_F16) return fail(SKR_E_DTYPE, "unknown dtype %d", dtype);
    if (desc->ndim < 1 || desc->ndim > SKR_MAX_DIMS) return fail(SKR_E_SHAPE, "ndim %d out of range", desc->ndim);
    if (desc->n_levels < 0 || desc->n_levels > SKR_MAX_LEVELS) return fail(SKR_E_SHAPE, "n_levels %d out of range", desc->n_levels);
    PyramidParams p;
    memset(&p, 0, sizeof(p));
    int64_t numel = 1;
    int masked = 0;
    for (int d = 0; d < desc->ndim; ++d) {
        if (desc->shape[d] < 1) return fail(SKR_E_SHAPE, "shape[%d] < 1", d);
        p.shape[d] = desc->shape[d];
        p.masked[d] = desc->masked[d] ? 1 : 0;
        masked += p.masked[d];
        numel *= desc->shape[d];
    }
    if (masked < 1 || masked > 2) return fail(SKR_E_UNSUPPORTED, "pyramid needs 1 or 2 resized axes, got %d", masked);
    if (!out) return fail(SKR_E_NULL, "null output");
    p.out = out; p.numel = numel; p.dtype = dtype; p.ndim = desc->ndim; p.n_levels = desc->n_levels; p.seed = desc->seed;
    for (int l = 0; l < desc->n_levels; ++l) {
        p.stream[l] = desc->levels[l].stream;
        p.buffer[l] = desc->levels[l].buffer;
        p.extent[l][0] = desc->levels[l].extent[0];
        p.extent[l][1] = desc->levels[l].extent[1];
        p.weight[l] = (float)desc->levels[l].weight;
        if (p.extent[l][0] < 1 || (masked == 2 && p.extent[l][1] < 1)) return fail(SKR_E_SHAPE, "level %d has an empty extent", l);
    }
    p.base_stream = desc->base_stream;
    p.base_buffer = desc->base_buffer;
    p.scratch = desc->scratch;
    p.moments = moments;
    p.m_axis[0] = p.m_axis[1] = -1;
    for (int d = 0, seen = 0; d < desc->ndim; ++d)
        if (p.masked[d] && seen < 2) p.m_axis[seen++] = d;
    bool narrow = numel < ((int64_t)1 << 31);
    for (int l = 0; l < desc->n_levels; ++l) {
        int64_t stride = 1;
        for (int d = desc->ndim - 1, seen = masked; d >= 0; --d) {
            p.lstride[l][d] = stride;
            stride *= p.masked[d] ? p.extent[l][--seen] : p.shape[d];
        }
        narrow = narrow && stride < ((int64_t)1 << 31);
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    auto pass = [&](int mode, const char* what) {
        p.mode = mode;
        if (narrow) pyramid_kernel<int32_t><<<grid_for(numel, 256), 256, 0, s>>>(p);
        else pyramid_kernel<int64_t><<<grid_for(numel, 256), 256, 0, s>>>(p);
        return check_launch(what);
    };
    for (int l = 0; l < desc->n_levels; ++l) {
        p.same_size[l] = p.extent[l][0] == p.shape[p.m_axis[0]] && (p.m_axis[1] < 0 || p.extent[l][1] == p.shape[p.m_axis[1]]);
        p.ratio[l][0] = (float)p.extent[l][0] / (float)p.shape[p.m_axis[0]];
        p.ratio[l][1] = p.m_axis[1] >= 0 ? (float)p.extent[l][1] / (float)p.shape[p.m_axis[1]] : 1.0f;
    }
    // The fused path: base draw and unit-sized levels in registers, coarse levels from grids (supplied, or drawn here
    // into `levels_scratch` by one launch), four elements per thread.
    float* levels_cursor = nullptr;  // first free float of levels_scratch after the level grids
    bool fused = p.scratch && narrow && (p.shape[desc->ndim - 1] & 3) == 0 && (reinterpret_cast<uintptr_t>(p.scratch) & 15u) == 0 &&
                 (!p.base_buffer || (reinterpret_cast<uintptr_t>(p.base_buffer) & 15u) == 0);
    if (desc->levels_scratch && (reinterpret_cast<uintptr_t>(desc->levels_scratch) & 15u) == 0) {
        // coarse levels without a buffer: all drawn into levels_scratch by one launch (both composition kernels then
        // interpolate grids instead of drawing every corner)
        LevelsFillParams fill;
        memset(&fill, 0, sizeof(fill));
        fill.seed = desc->seed;
        float* cursor = desc->levels_scratch;
        int64_t largest = 0;
        for (int l = 0; l < desc->n_levels; ++l) {
            if (p.weight[l] == 0.0f || p.same_size[l] || p.buffer[l]) continue;
            const int64_t count = p.lstride[l][0] * (p.masked[0] ? p.extent[l][0] : p.shape[0]);
            const int64_t used = cursor - desc->levels_scratch;
            if (count >= ((int64_t)1 << 31) || used + ((count + 3) & ~(int64_t)3) > desc->levels_scratch_floats) continue;  // stays an in-kernel draw
            fill.stream[l] = p.stream[l];
            fill.out[l] = cursor;
            fill.numel[l] = (int32_t)count;
            p.buffer[l] = cursor;
            cursor += (count + 3) & ~(int64_t)3;
            largest = count > largest ? count : largest;
        }
        if (largest > 0) {
            const unsigned blocks = grid_for((largest + 3) / 4, 256);
            levels_fill_kernel<<<dim3(blocks, (unsigned)desc->n_levels), 256, 0, s>>>(fill);
            int rc = check_launch("pyramid levels");
            if (rc) return rc;
        }
        levels_cursor = cursor;
    }
    for (int l = 0; l < desc->n_levels && fused; ++l) {
        if (p.weight[l] == 0.0f) continue;
        if (p.same_size[l]) fused = !p.buffer[l] || (reinterpret_cast<uintptr_t>(p.buffer[l]) & 15u) == 0;
        else fused = p.buffer[l] != nullptr;
    }
    if (fused) {
        // leading axes that are not resized only select a slice: collapse them so the kernel decomposes 2-3 indices
        int lead = 0;
        while (lead < desc->ndim && !p.masked[lead]) ++lead;
        bool trailing = lead + masked == desc->ndim;
        if (trailing) {
            int64_t slices = 1;
            for (int d = 0; d < lead; ++d) slices *= p.shape[d];
            const int nd = masked + 1;
            for (int l = 0; l < desc->n_levels; ++l) {
                const int64_t inner = masked == 2 ? p.extent[l][0] * p.extent[l][1] : p.extent[l][0];
                p.lstride[l][0] = inner;
                if (masked == 2) { p.lstride[l][1] = p.extent[l][1]; p.lstride[l][2] = 1; }
                else p.lstride[l][1] = 1;
            }
            const int64_t resized[2] = {p.shape[lead], masked == 2 ? p.shape[lead + 1] : 1};
            for (int k = 0; k < masked; ++k) { p.shape[1 + k] = resized[k]; p.masked[1 + k] = 1; }
            p.shape[0] = slices;
            p.masked[0] = 0;
            p.ndim = nd;
            p.m_axis[0] = 1;
            p.m_axis[1] = masked == 2 ? 2 : -1;
            // stretch the coarse levels to the unit's width first, when the scratch area has room after the grids
            {
                LevelsWidenParams widen;
                memset(&widen, 0, sizeof(widen));
                widen.width = (int32_t)p.shape[masked];
                float* cursor = levels_cursor;
                int64_t room = levels_cursor ? desc->levels_scratch_floats - (levels_cursor - desc->levels_scratch) : 0;
                int64_t largest = 0;
                for (int l = 0; l < desc->n_levels; ++l) {
                    if (p.weight[l] == 0.0f || p.same_size[l] || !p.buffer[l]) continue;
                    const int64_t rows = slices * (masked == 2 ? p.extent[l][0] : 1);
                    const int64_t count = rows * widen.width;
                    if (count > room || count >= ((int64_t)1 << 31)) continue;
                    widen.in[l] = p.buffer[l];
                    widen.out[l] = cursor;
                    widen.rows[l] = (int32_t)rows;
                    widen.in_width[l] = (int32_t)p.extent[l][masked - 1];
                    widen.ratio[l] = p.ratio[l][masked - 1];
                    p.wide[l] = cursor;
                    cursor += count;
                    room -= count;
                    largest = count > largest ? count : largest;
                }
                if (largest > 0) {
                    levels_widen_kernel<<<dim3(grid_for(largest / 4, 256), (unsigned)desc->n_levels), 256, 0, s>>>(widen);
                    int rc = check_launch("pyramid widen");
                    if (rc) return rc;
                }
            }
            {
                const int resident = launch_resident(p, masked, s);
                if (resident >= 0) return resident;
            }
            if (masked == 2) pyramid_compose_trailing_kernel<2><<<grid_for(numel / 4, 256), 256, 0, s>>>(p);
            else pyramid_compose_trailing_kernel<1><<<grid_for(numel / 4, 256), 256, 0, s>>>(p);
        } else {
            pyramid_compose4_kernel<<<grid_for(numel / 4, 256), 256, 0, s>>>(p);
        }
        int rc = check_launch("pyramid compose");
        if (rc) return rc;
        return skr_noise_scale(p.scratch, SKR_F32, out, dtype, numel, 1.0, nullptr, 0, moments, numel, 0.0, cuda_stream);
    }
    if (p.scratch) {
        int rc = pass(2, "pyramid compose");
        if (rc) return rc;
        return skr_noise_scale(p.scratch, SKR_F32, out, dtype, numel, 1.0, nullptr, 0, moments, numel, 0.0, cuda_stream);
    }
    int rc = pass(0, "pyramid moments");
    if (rc) return rc;
    return pass(1, "pyramid write");
}

int skr_colored_shape(void* spectrum, int32_t complex_dtype, const int64_t* dims, int32_t ndim, double exponent, void* cuda_stream) {
    using namespace skr;
    if (!spectrum || !dims) return fail(SKR_E_NULL, "null spectrum / dims");
    if (ndim < 1 || ndim > SKR_MAX_DIMS) return fail(SKR_E_SHAPE, "ndim %d out of range", ndim);
    if (complex_dtype != SKR_F32 && complex_dtype != SKR_F64) return fail(SKR_E_DTYPE, "spectrum must be complex64 or complex128");
    ShapeParams p;
    memset(&p, 0, sizeof(p));
    p.ndim = ndim;
    int64_t bins = 1;
    double sum = 0.0, rmax2 = 0.0;
    for (int d = 0; d < ndim; ++d) {
        if (dims[d] < 1) return fail(SKR_E_SHAPE, "dims[%d] < 1", d);
        p.dims[d] = dims[d];
        sum += (double)dims[d];
        bins *= d == ndim - 1 ? dims[d] / 2 + 1 : dims[d];
        const float fmax = (float)(dims[d] / 2) / (float)dims[d];  // largest |frequency| along the axis
        rmax2 += (double)(fmax * fmax);
    }
    p.bins = bins;
    if (complex_dtype == SKR_F32) p.spectrum = reinterpret_cast<float2*>(spectrum);
    else p.spectrum64 = reinterpret_cast<double2*>(spectrum);
    const double n_eff = sum / (double)ndim;
    p.eps_clip = (float)(0.5 / (n_eff > 4.0 ? n_eff : 4.0));
    p.r_max = sqrtf((float)rmax2);
    p.exponent_half_neg = (float)(-exponent / 2.0);
    const int64_t last_bins = dims[ndim - 1] / 2 + 1, rows = bins / last_bins;
    if (complex_dtype == SKR_F32 && rows >= 2048 && rows < ((int64_t)1 << 27) && last_bins >= 16 && last_bins < (1 << 20)) {
        colored_shape_rows_kernel<<<grid_for(rows * 32, 256), 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
            p, (int)rows, p.r_max > 0.0f ? 1.0f / p.r_max : 0.0f);
        return check_launch("colored shape");
    }
    if (bins < ((int64_t)1 << 31) - 256 * 148 * 8) colored_shape_kernel<int32_t><<<grid_for(bins, 256), 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(p);
    else colored_shape_kernel<int64_t><<<grid_for(bins, 256), 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(p);
    return check_launch("colored shape");
}

}  // extern "C"
