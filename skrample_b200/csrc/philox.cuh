// skrample_b200 - counter-based normal draws shared by the noise kernels and the fused step kernels.
//
// Philox4x32-10, key = 64-bit seed, counter = (element index / 4, stream id).  One block yields four standard
// normals (two Box-Muller pairs).  Every kernel that needs "the normal at element e of stream s" goes through
// these functions, so a noise tensor written by skr_noise_fill and the noise term drawn inside the step kernel
// are bit-identical.
#pragma once

#include <cstdint>

namespace skr {

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint64_t index, uint64_t stream) const {
        uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            // one 32x32 -> 64-bit multiply per product (mul.wide.u32 -> IMAD.WIDE.U32), halves split without arithmetic
            uint32_t lo0, hi0, lo1, hi1;
            asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(0xD2511F53u));
            asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(0xCD9E8D57u));
            c0 = hi1 ^ c1 ^ a;
            c1 = lo1;
            c2 = hi0 ^ c3 ^ b;
            c3 = lo0;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// (0, 1]; one fused multiply-add, written explicitly so that every translation unit (whatever its -fmad setting) gets
// the same bits
__device__ __forceinline__ float u01(uint32_t x) { return __fmaf_rn((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

// Box-Muller on the special-function unit.  A standard normal is needed to ~1e-5, not to the last bit (nothing pins the
// VALUES of device noise: torch's CUDA generator cannot be reproduced either; the contract is seed determinism, the
// moments, and that every kernel of this library draws the SAME value for the same (seed, stream, element)).  The
// full-precision logf + sqrtf + sincospif of the first version cost ~50 instructions per element and made the fill
// kernel issue-bound at 1.3 TB/s; this form is ~12 (MUFU.LG2 / SQRT / SIN / COS + a few FMAs), so a fill is bound by
// its HBM writes and a draw inside a step kernel hides under the step's loads.
//   -2 ln u:  lg2.approx has an ABSOLUTE error of ~2^-22 on [0.5, 2), which would swamp the result where u -> 1; there
//             the series 2t + t^2 + (2/3)t^3, t = 1 - u (exact: Sterbenz), is used instead (t < 2^-9: remainder < 2e-11).
//   sin/cos:  sin.approx / cos.approx of 2 pi v, absolute error ~4e-7 on [0, 2 pi].
// Resulting absolute error of a normal against exact arithmetic on the same uniforms: <= ~5e-6 (tests pin 1e-5).
__device__ __forceinline__ float neg2_log(float u) {
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u));
    const float t = 1.0f - u;
    const float series = __fmul_rn(t, __fmaf_rn(t, __fmaf_rn(t, 0.6666667f, 1.0f), 2.0f));
    return t < 0x1p-9f ? series : __fmul_rn(lg, -1.3862943611198906f);
}
__device__ __forceinline__ float sqrt_fast(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void sincos_turns(float v, float& s, float& c) {  // sin / cos of 2 pi v, v in (0, 1]
    const float a = __fmul_rn(v, 6.283185307179586f);
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(a));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(a));
}

// four standard normals from one Philox block (two Box-Muller pairs)
__device__ __forceinline__ void normal4(const uint4 r, float (&z)[4]) {
    const float r0 = sqrt_fast(neg2_log(u01(r.x)));
    const float r1 = sqrt_fast(neg2_log(u01(r.z)));
    float s0, c0, s1, c1;
    sincos_turns(u01(r.y), s0, c0);
    sincos_turns(u01(r.w), s1, c1);
    z[0] = __fmul_rn(r0, s0);
    z[1] = __fmul_rn(r0, c0);
    z[2] = __fmul_rn(r1, s1);
    z[3] = __fmul_rn(r1, c1);
}

// the normal at element `e` of stream (seed, stream)
__device__ __forceinline__ float normal_at(const Philox& ph, uint64_t e, uint64_t stream) {
    float z[4];
    normal4(ph(e >> 2, stream), z);
    const int lane = (int)(e & 3);
    return lane == 0 ? z[0] : lane == 1 ? z[1] : lane == 2 ? z[2] : z[3];
}

}  // namespace skr
