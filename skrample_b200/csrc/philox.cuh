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
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
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

__device__ __forceinline__ float u01(uint32_t x) { return (float)x * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }  // (0, 1]

// four standard normals from one Philox block (two Box-Muller pairs)
__device__ __forceinline__ void normal4(const uint4 r, float (&z)[4]) {
    const float r0 = sqrtf(-2.0f * logf(u01(r.x)));
    const float r1 = sqrtf(-2.0f * logf(u01(r.z)));
    float s0, c0, s1, c1;
    sincospif(2.0f * u01(r.y), &s0, &c0);
    sincospif(2.0f * u01(r.w), &s1, &c1);
    z[0] = r0 * s0;
    z[1] = r0 * c0;
    z[2] = r1 * s1;
    z[3] = r1 * c1;
}

// the normal at element `e` of stream (seed, stream)
__device__ __forceinline__ float normal_at(const Philox& ph, uint64_t e, uint64_t stream) {
    float z[4];
    normal4(ph(e >> 2, stream), z);
    const int lane = (int)(e & 3);
    return lane == 0 ? z[0] : lane == 1 ? z[1] : lane == 2 ? z[2] : z[3];
}

}  // namespace skr
