// Exhaustive check of the reciprocal-multiply division used by the step kernels (csrc/machine.cuh, Arith<float>::divr):
//
//     r  = RN(1/d)                       (once per divisor, on the host)
//     q0 = RN(a*r);  e = fma(-q0, d, a);  q = fma(e, r, q0)
//
// against IEEE division RN(a/d) for EVERY pair of fp32 significands (2^23 x 2^23).  With no overflow/underflow
// (the kernels guard the exponent range) all five operations commute exactly with scaling a or d by powers of
// two and with sign flips, so the significand pairs cover every guarded input.
//
// Build:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false tools/verify_divr.cu -o verify_divr
// Run:    ./verify_divr [first_d_mantissa [count [perturb_ulps]]]   (all 2^23 divisors: 45 s on a B200)
// Prints the number of mismatching pairs and the divisor significands that have any.  `perturb_ulps` moves the
// reciprocal off its correctly rounded value - the negative control showing the comparison can fail.
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__global__ void check(uint32_t d_first, uint32_t d_count, int perturb, unsigned long long* bad_pairs, uint32_t* bad_d, uint32_t* n_bad_d, uint32_t cap) {
    // one block per divisor, threads sweep the 2^23 dividends
    for (uint32_t di = blockIdx.x; di < d_count; di += gridDim.x) {
        const uint32_t dm = d_first + di;
        const float d = __uint_as_float(0x3f800000u | dm);
        const float r = __uint_as_float(__float_as_uint(__frcp_rn(d)) + perturb);
        uint32_t mine = 0;
        for (uint32_t am = threadIdx.x; am < (1u << 23); am += blockDim.x) {
            const float a = __uint_as_float(0x3f800000u | am);
            const float q0 = __fmul_rn(a, r);
            const float e = __fmaf_rn(-q0, d, a);
            const float q = __fmaf_rn(e, r, q0);
            mine += (__float_as_uint(q) != __float_as_uint(__fdiv_rn(a, d)));
        }
        // block reduce
        __shared__ uint32_t total;
        if (threadIdx.x == 0) total = 0;
        __syncthreads();
        if (mine) atomicAdd(&total, mine);
        __syncthreads();
        if (threadIdx.x == 0 && total) {
            atomicAdd(bad_pairs, (unsigned long long)total);
            const uint32_t slot = atomicAdd(n_bad_d, 1u);
            if (slot < cap) bad_d[slot] = dm;
        }
        __syncthreads();
    }
}

int main(int argc, char** argv) {
    const uint32_t first = argc > 1 ? (uint32_t)strtoul(argv[1], nullptr, 0) : 0u;
    const uint32_t count = argc > 2 ? (uint32_t)strtoul(argv[2], nullptr, 0) : (1u << 23) - first;
    const int perturb = argc > 3 ? atoi(argv[3]) : 0;
    const uint32_t cap = 1u << 16;
    unsigned long long* bad_pairs;
    uint32_t *bad_d, *n_bad_d;
    cudaMallocManaged(&bad_pairs, sizeof(*bad_pairs));
    cudaMallocManaged(&n_bad_d, sizeof(*n_bad_d));
    cudaMallocManaged(&bad_d, cap * sizeof(uint32_t));
    *bad_pairs = 0;
    *n_bad_d = 0;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    cudaEventRecord(t0);
    // chunks keep each launch short
    const uint32_t chunk = 1u << 16;
    for (uint32_t done = 0; done < count; done += chunk) {
        const uint32_t n = count - done < chunk ? count - done : chunk;
        check<<<148 * 8, 256>>>(first + done, n, perturb, bad_pairs, bad_d, n_bad_d, cap);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
    }
    cudaEventRecord(t1);
    cudaEventSynchronize(t1);
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    printf("divisor significands [%u, %u): %llu mismatching pairs, %u divisors affected, %.1f s\n", first, first + count,
           *bad_pairs, *n_bad_d, ms / 1000.0);
    for (uint32_t i = 0; i < *n_bad_d && i < 32; ++i) printf("  bad divisor significand 0x%06x\n", bad_d[i]);
    return *bad_pairs ? 1 : 0;
}
