// skrample_b200 - per-element arithmetic and operand fetch/store shared by the step kernels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "philox.cuh"


namespace skr {

constexpr int kThreads = 256;             // consumer threads per CTA
constexpr int kVec = 4;                   // elements per thread per tile
constexpr int kTile = kThreads * kVec;    // elements per tile
constexpr int kMaxStages = 8;

// ------------------------------------------------------------------------------------------
// individually rounded arithmetic

// General exponents of the SPC power mean go through double-precision pow, out of line: the code is large and only
// one uniform branch of the block kernel's generic instantiations and of the interpreter ever reaches it.
static __device__ __noinline__ float pow_general(float m, float f) { return (float)pow((double)m, (double)f); }
static __device__ __noinline__ double pow_general(double m, double f) { return pow(m, f); }

template <typename CT> struct Arith;
template <> struct Arith<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float spow(float x, float f) {
        // |x|^f * sign(x); torch's tensor-scalar pow has exact special cases (reference: common.py:187-190)
        const float m = fabsf(x);
        float r;
        if (f == 2.0f) r = __fmul_rn(m, m);
        else if (f == 3.0f) r = __fmul_rn(__fmul_rn(m, m), m);
        else if (f == 0.5f) r = __fsqrt_rn(m);
        else if (f == -0.5f) r = __fdiv_rn(1.0f, __fsqrt_rn(m));
        else if (f == -1.0f) r = __fdiv_rn(1.0f, m);
        else if (f == -2.0f) r = __fdiv_rn(1.0f, __fmul_rn(m, m));
        else r = pow_general(m, f);
        return x < 0.0f ? -r : r;
    }
};
template <> struct Arith<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double spow(double x, double f) {
        const double m = fabs(x);
        double r;
        if (f == 2.0) r = __dmul_rn(m, m);
        else if (f == 3.0) r = __dmul_rn(__dmul_rn(m, m), m);
        else if (f == 0.5) r = __dsqrt_rn(m);
        else if (f == -0.5) r = __ddiv_rn(1.0, __dsqrt_rn(m));
        else if (f == -1.0) r = __ddiv_rn(1.0, m);
        else if (f == -2.0) r = __ddiv_rn(1.0, __dmul_rn(m, m));
        else r = pow_general(m, f);
        return x < 0.0 ? -r : r;
    }
};

// Arithmetic policy of a block-kernel shape.  Exact (the default): every product, sum and quotient individually
// rounded in the reference's order - bit-identical to the reference's torch-CPU path.  Contracted (opt-in,
// skr_set_arithmetic(1) / SKR_ARITH=contracted): a*b + c is one fused multiply-add, a division by a uniform scalar is a
// multiplication by its reciprocal, the `0 +` of a sum's head is dropped - a few fp32 ulp per step (tests bound the
// trajectory at 1e-5 relative, the north star's own tolerance) for ~30% fewer instructions on the issue-bound steps.
template <typename CT, bool CONTRACT>
struct Policy : Arith<CT> {
    static constexpr bool contract = false;
    static __device__ __forceinline__ CT madd(CT acc, CT a, CT b) { return Arith<CT>::add(acc, Arith<CT>::mul(a, b)); }   // acc + a*b
    static __device__ __forceinline__ CT msub(CT acc, CT a, CT b) { return Arith<CT>::sub(acc, Arith<CT>::mul(a, b)); }   // acc - a*b
    static __device__ __forceinline__ CT head(CT a, CT b) { return Arith<CT>::add((CT)0, Arith<CT>::mul(a, b)); }          // 0 + a*b
};
template <>
struct Policy<float, true> : Arith<float> {
    static constexpr bool contract = true;
    static __device__ __forceinline__ float madd(float acc, float a, float b) { return __fmaf_rn(a, b, acc); }
    static __device__ __forceinline__ float msub(float acc, float a, float b) { return __fmaf_rn(-a, b, acc); }
    static __device__ __forceinline__ float head(float a, float b) { return __fmul_rn(a, b); }
};

// ------------------------------------------------------------------------------------------
// division by a grid-uniform scalar
//
// Every division of a solver step has a scalar divisor (alpha, r_k, the RK normaliser ...).  IEEE division costs
// ~12 instructions per element including a subroutine guard; with r = RN(1/d) from the host,
//
//     q0 = RN(a*r);  e = fma(-q0, d, a);  q = fma(e, r, q0)
//
// is the correctly rounded quotient whenever nothing under- or overflows (Markstein's correction step).  That is
// not taken on faith: tools/verify_divr.cu compares it with IEEE division for all 2^23 x 2^23 significand pairs on
// the GPU (result recorded in DESIGN.md).  The exponent range is guarded per group of elements: |d| in
// [2^-20, 2^20] is checked by the host (else `fast` is false), |a| in [2^-60, 2^60] here; zeros, subnormals,
// infinities and NaNs take the IEEE path out of line.

// Out of line and by value: the hot loop must not carry the IEEE sequence nor spill the operands to the stack.
static __device__ __noinline__ float4 div_ieee4(float4 a, float d) {
    return make_float4(__fdiv_rn(a.x, d), __fdiv_rn(a.y, d), __fdiv_rn(a.z, d), __fdiv_rn(a.w, d));
}

template <int V>
__device__ __forceinline__ void div_uniform(float (&a)[V], float d, float r, bool fast) {
#ifdef SKR_IEEE_DIV
    (void)r; (void)fast;
#pragma unroll
    for (int j = 0; j < V; ++j) a[j] = __fdiv_rn(a[j], d);
#else
    // r == 0: the host found THIS divisor outside the range it computes reciprocals for (the other divisions of the
    // launch keep the fast sequence)
    bool ok = fast && r != 0.0f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const float m = fabsf(a[j]);
        ok = ok && (m >= 0x1p-60f) && (m <= 0x1p60f);  // false for NaN
    }
    if (ok) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float q0 = __fmul_rn(a[j], r);
            const float e = __fmaf_rn(-q0, d, a[j]);
            a[j] = __fmaf_rn(e, r, q0);
        }
    } else if (r == 0.0f && fabsf(d) == __int_as_float(0x7f800000)) {
        // an infinite divisor - lambda = ln(alpha / sigma) of a flow schedule's first point, so the first `order` steps
        // of every flow-matching trajectory have one: x / +-inf IS x * +-0 in IEEE arithmetic (signed zeros for finite
        // x, NaN for inf / inf), so these steps need not leave the fast path either
        const float zero = copysignf(0.0f, d);
#pragma unroll
        for (int j = 0; j < V; ++j) a[j] = __fmul_rn(a[j], zero);
    } else {
        static_assert(V % 4 == 0, "elements per thread come in groups of four");
#pragma unroll
        for (int g = 0; g < V / 4; ++g) {
            const float4 q = div_ieee4(make_float4(a[4 * g], a[4 * g + 1], a[4 * g + 2], a[4 * g + 3]), d);
            a[4 * g] = q.x; a[4 * g + 1] = q.y; a[4 * g + 2] = q.z; a[4 * g + 3] = q.w;
        }
    }
#endif
}

// Contracted policy: one multiplication by the host's reciprocal (IEEE division when the divisor is outside the range
// the host computes reciprocals for).
template <int V>
__device__ __forceinline__ void div_reciprocal(float (&a)[V], float d, float r, bool fast) {
    if (fast && r != 0.0f) {
#pragma unroll
        for (int j = 0; j < V; ++j) a[j] = __fmul_rn(a[j], r);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) a[j] = __fdiv_rn(a[j], d);
    }
}

template <int V>
__device__ __forceinline__ void div_uniform(double (&a)[V], double d, double, bool) {
#pragma unroll
    for (int j = 0; j < V; ++j) a[j] = __ddiv_rn(a[j], d);
}

// Host side: r = RN(1/d) in the compute type, or 0 when d is outside the guarded range.
static inline float uniform_reciprocal(float d, bool*) {
    const float m = d < 0 ? -d : d;
    if (!(m >= 0x1p-20f && m <= 0x1p20f)) return 0.0f;  // this divisor takes IEEE division (or the +-inf shortcut)
    return 1.0f / d;  // IEEE division on the host: correctly rounded
}
static inline double uniform_reciprocal(double, bool*) { return 0.0; }

// ------------------------------------------------------------------------------------------
// operand fetch / result store

// Staged path: this thread's 4 consecutive elements of input `i` from the shared-memory tile.
template <typename CT>
__device__ __forceinline__ void fetch_staged(const unsigned char* stage, uint32_t off, int dtype, int tid, CT (&v)[kVec]) {
    const unsigned char* base = stage + off;
    switch (dtype) {
        case SKR_F32: {
            const float4 q = *reinterpret_cast<const float4*>(base + tid * 16);
            v[0] = (CT)q.x; v[1] = (CT)q.y; v[2] = (CT)q.z; v[3] = (CT)q.w;
        } break;
        case SKR_BF16: {
            const uint2 q = *reinterpret_cast<const uint2*>(base + tid * 8);
            v[0] = (CT)__uint_as_float(q.x << 16); v[1] = (CT)__uint_as_float(q.x & 0xffff0000u);
            v[2] = (CT)__uint_as_float(q.y << 16); v[3] = (CT)__uint_as_float(q.y & 0xffff0000u);
        } break;
        case SKR_F16: {
            const uint2 q = *reinterpret_cast<const uint2*>(base + tid * 8);
            const __half2 lo = *reinterpret_cast<const __half2*>(&q.x);
            const __half2 hi = *reinterpret_cast<const __half2*>(&q.y);
            const float2 a = __half22float2(lo), b = __half22float2(hi);
            v[0] = (CT)a.x; v[1] = (CT)a.y; v[2] = (CT)b.x; v[3] = (CT)b.y;
        } break;
        default: {  // SKR_F64 - only reachable in the fp64-compute instantiation
            if constexpr (sizeof(CT) == 8) {
                const double2 q0 = *reinterpret_cast<const double2*>(base + tid * 32);
                const double2 q1 = *reinterpret_cast<const double2*>(base + tid * 32 + 16);
                v[0] = (CT)q0.x; v[1] = (CT)q0.y; v[2] = (CT)q1.x; v[3] = (CT)q1.y;
            }
        } break;
    }
}

// Guarded path: element-wise loads straight from global memory (tail tile / unaligned tensors).
template <typename CT>
__device__ __forceinline__ void fetch_direct(const void* ptr, int dtype, int64_t first, int64_t numel, CT (&v)[kVec]) {
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
        const int64_t e = first + j;
        CT x = (CT)0;
        if (e < numel) {
            switch (dtype) {
                case SKR_F32: x = (CT) reinterpret_cast<const float*>(ptr)[e]; break;
                case SKR_BF16: x = (CT)__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ptr)[e]); break;
                case SKR_F16: x = (CT)__half2float(reinterpret_cast<const __half*>(ptr)[e]); break;
                default: x = (CT) reinterpret_cast<const double*>(ptr)[e]; break;
            }
        }
        v[j] = x;
    }
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&p);
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    const __half2 p = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&p);
}
// double -> bf16/f16 goes through one rounding only (to-odd trick is not needed: these outputs
// are produced from fp32 compute in every supported program; fp64 compute with 16-bit storage
// rounds double->float->half, documented in DESIGN.md).

template <typename CT, bool DIRECT>
__device__ __forceinline__ void store_vec(void* ptr, int dtype, int64_t first, int64_t numel, const CT (&v)[kVec]) {
    if constexpr (!DIRECT) {
        switch (dtype) {
            case SKR_F32:
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(ptr) + first) =
                    make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
                break;
            case SKR_BF16:
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(ptr) + first) =
                    make_uint2(pack_bf16((float)v[0], (float)v[1]), pack_bf16((float)v[2], (float)v[3]));
                break;
            case SKR_F16:
                *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ptr) + first) =
                    make_uint2(pack_f16((float)v[0], (float)v[1]), pack_f16((float)v[2], (float)v[3]));
                break;
            default: {
                double* p = reinterpret_cast<double*>(ptr) + first;
                *reinterpret_cast<double2*>(p) = make_double2((double)v[0], (double)v[1]);
                *reinterpret_cast<double2*>(p + 2) = make_double2((double)v[2], (double)v[3]);
            } break;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kVec; ++j) {
            const int64_t e = first + j;
            if (e < numel) {
                switch (dtype) {
                    case SKR_F32: reinterpret_cast<float*>(ptr)[e] = (float)v[j]; break;
                    case SKR_BF16: reinterpret_cast<__nv_bfloat16*>(ptr)[e] = __float2bfloat16_rn((float)v[j]); break;
                    case SKR_F16: reinterpret_cast<__half*>(ptr)[e] = __float2half_rn((float)v[j]); break;
                    default: reinterpret_cast<double*>(ptr)[e] = (double)v[j]; break;
                }
            }
        }
    }
}


// Device copy of skr_philox (kernel-parameter constant bank).
struct KPhilox {
    uint64_t seed[SKR_MAX_PHILOX_ITEMS];
    uint64_t stream[SKR_MAX_PHILOX_ITEMS];
    int64_t item_numel;
    int32_t n_items;
    int32_t aligned;  // bit 0: item_numel % 4 == 0 (four consecutive elements share one Philox block);
                      // bit 1: numel < 2^31 (32-bit index arithmetic);
                      // bits 8-9: round every normal to bf16 (1) / fp16 (2), like the tensor skr_noise_fill writes
    int64_t offset_inner;  // Offset noise: > 0 = elements of an item that share one offset draw (skr_philox)
    float offset_scale;
    int32_t reserved;
};

// The key tables of a launch that draws noise in the kernel travel as a second kernel parameter, so the
// instantiations that read their noise from memory carry no table at all (their parameter block stays ~3 KB).
template <bool ON>
struct PhiloxKeys {};
template <>
struct PhiloxKeys<true> {
    KPhilox table[SKR_MAX_PHILOX];
};

__device__ __forceinline__ float round_as_stored(float z, int mode) {
    if (mode == 1) return __bfloat162float(__float2bfloat16_rn(z));
    if (mode == 2) return __half2float(__float2half_rn(z));
    return z;
}

// Offset noise: the shared draw of the run an element belongs to, added with two individually rounded operations (the
// fill kernel of the noise unit, compiled with contraction on, spells the same two out).  Out of line: plain Random
// draws - the hot case - pay one uniform branch for it.
static __device__ __noinline__ float4 add_offsets(uint64_t seed, uint64_t stream, int64_t inner, float scale, int64_t local, float4 z) {
    const Philox ph(seed);
    const int64_t row = local / inner;
    if (local + 3 < (row + 1) * inner) {
        const float shift = __fmul_rn(normal_at(ph, (uint64_t)row, stream), scale);
        z.x = __fadd_rn(z.x, shift); z.y = __fadd_rn(z.y, shift); z.z = __fadd_rn(z.z, shift); z.w = __fadd_rn(z.w, shift);
    } else {
        z.x = __fadd_rn(z.x, __fmul_rn(normal_at(ph, (uint64_t)(local / inner), stream), scale));
        z.y = __fadd_rn(z.y, __fmul_rn(normal_at(ph, (uint64_t)((local + 1) / inner), stream), scale));
        z.z = __fadd_rn(z.z, __fmul_rn(normal_at(ph, (uint64_t)((local + 2) / inner), stream), scale));
        z.w = __fadd_rn(z.w, __fmul_rn(normal_at(ph, (uint64_t)((local + 3) / inner), stream), scale));
    }
    return z;
}

// Four normals of the virtual noise tensor starting at element e (e % 4 == 0, item_numel % 4 == 0).  OFFSETS = false
// compiles the Offset term out (the pinned step shapes: launches whose draws carry offsets take the generic shape).
template <bool OFFSETS>
__device__ __forceinline__ float4 draw_group(const KPhilox& d, int64_t e) {
    int64_t item, local;
    if (d.aligned & 2) {
        const uint32_t i32 = (uint32_t)e / (uint32_t)d.item_numel;
        item = i32;
        local = (uint32_t)e - i32 * (uint32_t)d.item_numel;
    } else {
        item = e / d.item_numel;
        local = e - item * d.item_numel;
    }
    float z[4];
    normal4(Philox(d.seed[item])((uint64_t)local >> 2, d.stream[item]), z);
    if (OFFSETS && d.offset_inner > 0) {
        const float4 shifted = add_offsets(d.seed[item], d.stream[item] + 1, d.offset_inner, d.offset_scale, local, make_float4(z[0], z[1], z[2], z[3]));
        z[0] = shifted.x; z[1] = shifted.y; z[2] = shifted.z; z[3] = shifted.w;
    }
    const int mode = (d.aligned >> 8) & 3;
    if (mode) {
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = round_as_stored(z[j], mode);
    }
    return make_float4(z[0], z[1], z[2], z[3]);
}

template <bool OFFSETS>
__device__ __forceinline__ float draw_single(const KPhilox& d, int64_t e) {
    const int64_t item = e / d.item_numel;
    const int64_t local = e - item * d.item_numel;
    float z = normal_at(Philox(d.seed[item]), (uint64_t)local, d.stream[item]);
    if (OFFSETS && d.offset_inner > 0)
        z = __fadd_rn(z, __fmul_rn(normal_at(Philox(d.seed[item]), (uint64_t)(local / d.offset_inner), d.stream[item] + 1), d.offset_scale));
    return round_as_stored(z, (d.aligned >> 8) & 3);
}

// V consecutive elements starting at global element `first` (a multiple of 4) of the virtual noise tensor.
template <typename CT, int V, bool OFFSETS = true>
__device__ __forceinline__ void draw_normals(const KPhilox& d, int64_t first, int64_t numel, CT (&v)[V]) {
    if (d.aligned & 1) {
#pragma unroll
        for (int g = 0; g < V / 4; ++g) {
            const int64_t e = first + 4 * g;
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < numel) z = draw_group<OFFSETS>(d, e);
            v[4 * g] = (CT)z.x; v[4 * g + 1] = (CT)z.y; v[4 * g + 2] = (CT)z.z; v[4 * g + 3] = (CT)z.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const int64_t e = first + j;
            v[j] = (CT)(e < numel ? draw_single<OFFSETS>(d, e) : 0.f);
        }
    }
}

static inline void fill_kphilox(KPhilox* out, const skr_philox* in, int count) {
    for (int i = 0; i < count; ++i) {
        const int n = in[i].n_items < SKR_MAX_PHILOX_ITEMS ? in[i].n_items : SKR_MAX_PHILOX_ITEMS;
        for (int j = 0; j < n; ++j) { out[i].seed[j] = in[i].seed[j]; out[i].stream[j] = in[i].stream[j]; }
        out[i].item_numel = in[i].item_numel;
        out[i].n_items = in[i].n_items;
        out[i].offset_inner = in[i].offset_scale != 0.0f ? in[i].offset_inner : 0;
        out[i].offset_scale = in[i].offset_scale;
        out[i].reserved = 0;
        const int round_to = in[i].dtype == SKR_BF16 ? 1 : in[i].dtype == SKR_F16 ? 2 : 0;
        out[i].aligned = ((in[i].item_numel % 4) == 0 ? 1 : 0) |
                         (in[i].item_numel * in[i].n_items < (int64_t)0x7fffffff ? 2 : 0) | (round_to << 8);
    }
}

}  // namespace skr
