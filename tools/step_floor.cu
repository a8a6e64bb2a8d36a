// Where does the time of a small (launch-bound) solver step go?  A stand-alone timeline probe for BASELINE.json configs[1]
// (UniPC-3 SDE, 8x4x128x128 bf16: 512 tiles of 1024 elements, 8 input tensors = 24 KB per tile, 3 outputs = 10 KB).
//
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/step_floor tools/step_floor.cu && /tmp/step_floor
//
// It replays chains of 400 launches from a CUDA graph, every launch with the programmatic-dependent-launch attribute
// and the same grid / block / shared-memory shape as skr::block_kernel on that step, rotating over 17 sets of buffers
// (working set > 2x L2, like bench.py), and prints the time per launch of
//   empty       griddepcontrol.launch_dependents + wait, nothing else          -> launch + scheduling floor
//   load        + one 1-D TMA bulk copy per input tensor into shared memory     -> + HBM read phase
//   load+store  + the three outputs written from registers (no arithmetic)      -> + HBM write phase
// The product kernel on the same step is bench.py's `kernel_only.ms_per_step`; the difference to load+store is the
// arithmetic, the staging handshake and the register traffic.  Development aid: results are quoted in profiles/.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } \
    } while (0)

constexpr int kThreads = 288, kConsumers = 256, kTile = 1024, kTiles = 512;
constexpr int kInputs = 8, kOutputs = 3;
__host__ __device__ constexpr int in_bytes(int i) { return i < 4 ? 2 : 4; }  // x, out, previous noise, noise (bf16); previous sample, 3 x-hat (fp32)
constexpr int kOutBytes[kOutputs] = {4, 4, 2};                // corrected sample, x-hat (fp32), final (bf16)

struct Step {
    const unsigned char* in[kInputs];
    unsigned char* out[kOutputs];
    int mode;  // 0 empty, 1 load, 2 load + store
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kThreads, 4) probe(const __grid_constant__ Step s) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (s.mode > 0 && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (s.mode == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == kConsumers / 32) {  // the producer warp: one lane per input tensor
        uint32_t total = 0;
        for (int i = 0; i < kInputs; ++i) total += kTile * in_bytes(i);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(total) : "memory");
        __syncwarp();
        if (lane < kInputs) {
            uint32_t off = 0;
            for (int i = 0; i < lane; ++i) off += kTile * in_bytes(i);
            const uint32_t bytes = kTile * in_bytes(lane);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + off)),
                         "l"(s.in[lane] + (size_t)blockIdx.x * bytes), "r"(bytes), "r"(smem_u32(&bar))
                         : "memory");
        }
        return;
    }
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(&bar)),
        "r"(0)
        : "memory");
    if (s.mode < 2) return;
    const float4 v = *reinterpret_cast<const float4*>(smem + 8 * kTile + threadIdx.x * 16);  // something staged, so the stores depend on the loads
    const size_t first = (size_t)blockIdx.x * kTile + threadIdx.x * 4;
    *reinterpret_cast<float4*>(s.out[0] + first * 4) = v;
    *reinterpret_cast<float4*>(s.out[1] + first * 4) = v;
    *reinterpret_cast<uint2*>(s.out[2] + first * 2) = make_uint2(__float_as_uint(v.x), __float_as_uint(v.y));
}

int main() {
    constexpr int kSets = 17, kLaunches = 400;
    const size_t numel = (size_t)kTiles * kTile;
    std::vector<Step> sets(kSets);
    for (Step& s : sets) {
        for (int i = 0; i < kInputs; ++i) {
            unsigned char* p;
            CHECK(cudaMalloc(&p, numel * in_bytes(i)));
            CHECK(cudaMemset(p, 0, numel * in_bytes(i)));
            s.in[i] = p;
        }
        for (int i = 0; i < kOutputs; ++i) CHECK(cudaMalloc(&s.out[i], numel * kOutBytes[i]));
    }
    int smem = 0;
    for (int i = 0; i < kInputs; ++i) smem += kTile * in_bytes(i);
    CHECK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * smem));
    cudaStream_t stream;
    CHECK(cudaStreamCreate(&stream));
    const char* names[3] = {"empty", "load", "load+store"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int pdl = 1; pdl >= 0; --pdl) {
            cudaGraph_t graph;
            cudaGraphExec_t exec;
            CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            for (int k = 0; k < kLaunches; ++k) {
                Step s = sets[k % kSets];
                s.mode = mode;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(kTiles);
                cfg.blockDim = dim3(kThreads);
                cfg.dynamicSmemBytes = 2 * smem;  // two stages, like the product kernel on this step
                cfg.stream = stream;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr;
                cfg.numAttrs = pdl;
                CHECK(cudaLaunchKernelEx(&cfg, probe, s));
            }
            CHECK(cudaStreamEndCapture(stream, &graph));
            CHECK(cudaGraphInstantiate(&exec, graph, 0));
            cudaEvent_t a, b;
            CHECK(cudaEventCreate(&a));
            CHECK(cudaEventCreate(&b));
            for (int w = 0; w < 5; ++w) CHECK(cudaGraphLaunch(exec, stream));
            CHECK(cudaEventRecord(a, stream));
            const int reps = 50;
            for (int r = 0; r < reps; ++r) CHECK(cudaGraphLaunch(exec, stream));
            CHECK(cudaEventRecord(b, stream));
            CHECK(cudaStreamSynchronize(stream));
            float ms = 0;
            CHECK(cudaEventElapsedTime(&ms, a, b));
            printf("%-11s pdl=%d  %.3f us per launch\n", names[mode], pdl, ms * 1e3 / (reps * kLaunches));
            CHECK(cudaGraphExecDestroy(exec));
            CHECK(cudaGraphDestroy(graph));
        }
    }
    double bytes_in = 0, bytes_out = 0;
    for (int i = 0; i < kInputs; ++i) bytes_in += (double)numel * in_bytes(i);
    for (int i = 0; i < kOutputs; ++i) bytes_out += (double)numel * kOutBytes[i];
    printf("bytes per launch: %.2f MB read, %.2f MB written\n", bytes_in / 1e6, bytes_out / 1e6);
    return 0;
}
