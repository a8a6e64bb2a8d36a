// skrample_b200 - fused solver-step kernel for sm_100a (B200).
//
// One launch executes a whole step program (include/skrample_b200.h) over a latent batch:
//   * a persistent grid (one or two CTAs per SM) walks 1024-element tiles;
//   * the inputs of a tile are staged into shared memory by 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), several tiles ahead, so the memory-level
//     parallelism that saturates HBM3e does not depend on register count or occupancy;
//   * 256 threads interpret the program with eight named per-element registers held in
//     real registers (4 elements per thread), reading operands from the staged tile;
//   * every arithmetic op is an individually rounded IEEE op (__fmul_rn / __fadd_rn /
//     __fdiv_rn, no FMA contraction) in the reference's evaluation order, so fp32/fp64
//     results are bit-identical to the reference torch-CPU path;
//   * results are rounded once to the storage dtype and written with 128-bit stores.
//
// The op loop is driven from kernel parameters (constant bank, uniform across the grid):
// all control flow is warp-uniform.  Tail elements and unaligned tensors run through a
// guarded element-wise path of the same interpreter.

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>

#include "../../include/skrample_b200.h"
#include "common.cuh"
#include "machine.cuh"
#include "block_kernel.cuh"
#include "block_launch.cuh"
#include "host.cuh"

namespace skr {

template <typename CT>
struct KOp {
    uint8_t code, a, b, pad;
    int16_t src, dst;
    CT c[4];
};

template <typename CT>
struct KProgram {
    int32_t n_ops, n_inputs, n_outputs, stages;
    int64_t numel;
    uint32_t stage_bytes;   // bytes of one staged tile (all inputs)
    uint32_t use_tma;       // 0: every tile goes through the guarded element-wise path
    KOp<CT> ops[SKR_MAX_OPS];
    const void* in_ptr[SKR_MAX_INPUTS];
    void* out_ptr[SKR_MAX_OUTPUTS];
    uint32_t in_off[SKR_MAX_INPUTS];  // byte offset of input i inside a stage
    uint8_t in_dtype[SKR_MAX_INPUTS];
    uint8_t out_dtype[SKR_MAX_OUTPUTS];
    KPhilox philox[SKR_MAX_PHILOX];
};

// ------------------------------------------------------------------------------------------
// register file access by run-time id (uniform switch, the registers themselves stay static)

#define SKR_REG_SWITCH(id, STMT)                                                                      \
    switch (id) {                                                                                     \
        case 0: { auto& reg = r0; STMT; } break;                                                      \
        case 1: { auto& reg = r1; STMT; } break;                                                      \
        case 2: { auto& reg = r2; STMT; } break;                                                      \
        case 3: { auto& reg = r3; STMT; } break;                                                      \
        case 4: { auto& reg = r4; STMT; } break;                                                      \
        case 5: { auto& reg = r5; STMT; } break;                                                      \
        case 6: { auto& reg = r6; STMT; } break;                                                      \
        default: { auto& reg = r7; STMT; } break;                                                     \
    }

template <typename CT, bool DIRECT>
__device__ __forceinline__ void run_tile(const KProgram<CT>& prog, int64_t tile, const unsigned char* stage, int tid) {
    using Ar = Arith<CT>;
    // X, P, B, A, S, R, T, U
    CT r0[kVec], r1[kVec], r2[kVec], r3[kVec], r4[kVec], r5[kVec], r6[kVec], r7[kVec];
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
        r0[j] = r1[j] = r2[j] = r3[j] = r4[j] = r5[j] = r6[j] = r7[j] = (CT)0;
    }
    CT(&rX)[kVec] = r0; CT(&rP)[kVec] = r1; CT(&rB)[kVec] = r2; CT(&rA)[kVec] = r3;
    CT(&rS)[kVec] = r4; CT(&rR)[kVec] = r5; CT(&rT)[kVec] = r6; CT(&rU)[kVec] = r7;

    const int64_t first = tile * kTile + (int64_t)tid * kVec;
    const int64_t numel = prog.numel;
    const int n_ops = prog.n_ops;

    for (int pc = 0; pc < n_ops; ++pc) {
        const KOp<CT>& op = prog.ops[pc];
        const int code = op.code, a = op.a, b = op.b, src = op.src;
        const CT c0 = op.c[0], c1 = op.c[1], c2 = op.c[2], c3 = op.c[3];

        CT in[kVec];
        if (code == SKR_OP_FWD && (b & 2)) {
            draw_normals<CT, kVec>(prog.philox[src], first, numel, in);
        } else if (src >= 0) {
            if constexpr (DIRECT) fetch_direct<CT>(prog.in_ptr[src], prog.in_dtype[src], first, numel, in);
            else fetch_staged<CT>(stage, prog.in_off[src], prog.in_dtype[src], tid, in);
        } else {
#pragma unroll
            for (int j = 0; j < kVec; ++j) in[j] = (CT)0;
        }

        switch (code) {
            case SKR_OP_LOAD: {
                if (b & 1) {
#pragma unroll
                    for (int j = 0; j < kVec; ++j) in[j] = -in[j];
                }
                SKR_REG_SWITCH(a, _Pragma("unroll") for (int j = 0; j < kVec; ++j) reg[j] = in[j]);
            } break;
            case SKR_OP_MOV: {
                CT t[kVec];
                SKR_REG_SWITCH(b, _Pragma("unroll") for (int j = 0; j < kVec; ++j) t[j] = reg[j]);
                SKR_REG_SWITCH(a, _Pragma("unroll") for (int j = 0; j < kVec; ++j) reg[j] = t[j]);
            } break;
            case SKR_OP_STORE: {
                CT t[kVec];
                SKR_REG_SWITCH(a, _Pragma("unroll") for (int j = 0; j < kVec; ++j) t[j] = reg[j]);
                store_vec<CT, DIRECT>(prog.out_ptr[op.dst], prog.out_dtype[op.dst], first, numel, t);
            } break;
            case SKR_OP_CONV: {
#pragma unroll
                for (int j = 0; j < kVec; ++j) {
                    const CT y = (b == 0) ? in[j] : rP[j];
                    CT v;
                    if (a & SKR_CONV_USE_X) {
                        const CT lhs = (a & SKR_CONV_MUL_X) ? Ar::mul(c0, rX[j]) : rX[j];
                        const CT rhs = (a & SKR_CONV_MUL_Y) ? Ar::mul(c1, y) : y;
                        v = Ar::sub(lhs, rhs);
                    } else {
                        v = (a & SKR_CONV_MUL_Y) ? Ar::mul(y, c1) : y;
                    }
                    rP[j] = (a & SKR_CONV_DIV) ? Ar::div(v, c2) : v;
                }
            } break;
            case SKR_OP_ACC0:
            case SKR_OP_ACC: {
                if (a != 0) {
                    SKR_REG_SWITCH(a - 1, _Pragma("unroll") for (int j = 0; j < kVec; ++j) in[j] = reg[j]);
                }
#pragma unroll
                for (int j = 0; j < kVec; ++j) {
                    const CT base = (code == SKR_OP_ACC0) ? (CT)0 : rA[j];
                    rA[j] = Ar::add(base, Ar::mul(in[j], c0));
                }
            } break;
            case SKR_OP_DIVA: {
#pragma unroll
                for (int j = 0; j < kVec; ++j) rA[j] = Ar::div(rA[j], c0);
            } break;
            case SKR_OP_UNI: {
#pragma unroll
                for (int j = 0; j < kVec; ++j) {
                    const CT term = Ar::mul(Ar::div(Ar::sub(in[j], rB[j]), c0), c1);
                    rA[j] = Ar::add(a ? (CT)0 : rA[j], term);
                }
            } break;
            case SKR_OP_UNIC: {
#pragma unroll
                for (int j = 0; j < kVec; ++j) {
                    const CT term = Ar::mul(Ar::sub(rP[j], rB[j]), c1);
                    rA[j] = Ar::add(a ? (CT)0 : rA[j], term);
                }
            } break;
            case SKR_OP_ADDB: {
#pragma unroll
                for (int j = 0; j < kVec; ++j) rA[j] = Ar::add(rB[j], a ? (CT)0 : rA[j]);
            } break;
            case SKR_OP_DPM2: {
#pragma unroll
                for (int j = 0; j < kVec; ++j)
                    rA[j] = Ar::add(rB[j], Ar::mul(c1, Ar::mul(c0, Ar::sub(rB[j], in[j]))));
            } break;
            case SKR_OP_DPM3A: {
#pragma unroll
                for (int j = 0; j < kVec; ++j) {
                    rT[j] = in[j];
                    rU[j] = Ar::mul(c0, Ar::sub(rB[j], in[j]));
                }
            } break;
            case SKR_OP_DPM3B: {
#pragma unroll
                for (int j = 0; j < kVec; ++j) {
                    const CT d11 = Ar::mul(c0, Ar::sub(rT[j], in[j]));
                    const CT d10 = rU[j];
                    const CT d = Ar::sub(d10, d11);
                    rT[j] = Ar::add(d10, Ar::mul(c1, d));
                    rU[j] = Ar::mul(c2, d);
                }
            } break;
            case SKR_OP_DPM3C: {
#pragma unroll
                for (int j = 0; j < kVec; ++j)
                    rA[j] = Ar::add(Ar::add(rB[j], Ar::mul(c0, rT[j])), Ar::mul(c1, rU[j]));
            } break;
            case SKR_OP_FWD: {
                CT p[kVec];
                SKR_REG_SWITCH(a, _Pragma("unroll") for (int j = 0; j < kVec; ++j) p[j] = reg[j]);
#pragma unroll
                for (int j = 0; j < kVec; ++j) {
                    CT v = Ar::add((CT)0, Ar::mul(rX[j], c0));
                    v = Ar::add(v, Ar::mul(p[j], c1));
                    if (b & 3) v = Ar::add(v, Ar::mul(in[j], c2));
                    rR[j] = v;
                }
            } break;
            case SKR_OP_BACK: {
#pragma unroll
                for (int j = 0; j < kVec; ++j) {
                    CT v = Ar::sub(rR[j], Ar::mul(rX[j], c0));
                    if (b & 1) v = Ar::sub(v, Ar::mul(in[j], c2));
                    rP[j] = Ar::div(v, c1);
                }
            } break;
            case SKR_OP_BLEND: {
#pragma unroll
                for (int j = 0; j < kVec; ++j) {
                    if (a == 0) {
                        rX[j] = Ar::add(Ar::mul(rS[j], c0), Ar::mul(rR[j], c1));
                    } else {
                        const CT mixed = Ar::add(Ar::mul(Ar::spow(rS[j], c2), c0), Ar::mul(Ar::spow(rR[j], c2), c1));
                        rX[j] = Ar::spow(mixed, c3);
                    }
                }
            } break;
            case SKR_OP_AXPBY: {
#pragma unroll
                for (int j = 0; j < kVec; ++j) {
                    if (a == 0) rR[j] = Ar::add(Ar::mul(rX[j], c0), Ar::mul(in[j], c1));
                    else rR[j] = Ar::div(Ar::sub(rX[j], Ar::mul(in[j], c0)), c1);
                }
            } break;
            default: break;
        }
    }
}

// ------------------------------------------------------------------------------------------
// the kernel

template <typename CT>
__global__ void __launch_bounds__(kThreads) step_kernel(const __grid_constant__ KProgram<CT> prog) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];

    const int tid = threadIdx.x;
    const int64_t numel = prog.numel;
    const int64_t n_full = prog.use_tma ? numel / kTile : 0;
    const int64_t n_tiles = (numel + kTile - 1) / kTile;
    const int stages = prog.stages;
    const uint32_t stage_bytes = prog.stage_bytes;

    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int64_t mine = n_full > (int64_t)blockIdx.x ? (n_full - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (mine > 0) {
        if (tid == 0) {
            for (int s = 0; s < stages; ++s) mbar_init(&full_bar[s], 1);
            fence_barrier_init();
        }
        __syncthreads();

        auto issue = [&](int64_t k) {
            const int s = (int)(k % stages);
            const int64_t tile = blockIdx.x + k * (int64_t)gridDim.x;
            unsigned char* dst = smem + (size_t)s * stage_bytes;
            mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
            for (int i = 0; i < prog.n_inputs; ++i) {
                const uint32_t esize = dtype_size(prog.in_dtype[i]);
                const unsigned char* srcp = reinterpret_cast<const unsigned char*>(prog.in_ptr[i]) + (size_t)tile * kTile * esize;
                tma_load_1d(dst + prog.in_off[i], srcp, kTile * esize, &full_bar[s]);
            }
        };

        if (tid == 0) {
            const int64_t ahead = mine < stages - 1 ? mine : stages - 1;
            for (int64_t k = 0; k < ahead; ++k) issue(k);
        }
        for (int64_t k = 0; k < mine; ++k) {
            if (tid == 0 && k + stages - 1 < mine) issue(k + stages - 1);
            const int s = (int)(k % stages);
            mbar_wait(&full_bar[s], (uint32_t)((k / stages) & 1));
            run_tile<CT, false>(prog, blockIdx.x + k * (int64_t)gridDim.x, smem + (size_t)s * stage_bytes, tid);
            __syncthreads();  // everyone is done with stage s before it is refilled
        }
    }

    // remaining tiles (the ragged tail, or everything when TMA staging is off): guarded path
    for (int64_t tile = n_full + blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        run_tile<CT, true>(prog, tile, nullptr, tid);
    }
}

// ------------------------------------------------------------------------------------------
// host side
//
// Re-entrant by construction: the only process-wide state is (a) atomic launch counters, (b) the per-device
// attribute table filled under std::call_once, (c) the development switches read from the environment once
// (skr_reload_env re-reads them) and (d) per-instantiation "attribute set" bit masks (atomic).  Error text is
// thread-local; descriptors live on the caller's stack or inside an immutable skr_plan.

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};
static std::atomic<int64_t> g_launches_kind[3];  // [0] structured block kernel, [1] interpreter, [2] noise kernels

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

// Development switches (tests and A/B tooling): read once, re-read by skr_reload_env().
struct Switches {
    std::atomic<int> force_interp{0}, no_pinned{0}, in_mode{-1}, stages{0}, ctas{0};
    std::atomic<int> arithmetic{0};  // 0 exact, 1 contracted (skr_set_arithmetic; initial value from SKR_ARITH)
};
static Switches g_switches;
static std::once_flag g_switches_once;

static int env_raw(const char* name, int fallback) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : fallback;
}
static void load_switches() {
    g_switches.force_interp = env_raw("SKR_FORCE_INTERP", 0);
    g_switches.no_pinned = env_raw("SKR_NO_PINNED", 0);
    g_switches.in_mode = env_raw("SKR_IN_MODE", -1);
    g_switches.stages = env_raw("SKR_STAGES", 0);
    g_switches.ctas = env_raw("SKR_CTAS", 0);
    const char* arith = getenv("SKR_ARITH");
    g_switches.arithmetic = (arith && (arith[0] == 'c' || arith[0] == '1')) ? 1 : 0;
}
static const Switches& switches() {
    std::call_once(g_switches_once, load_switches);
    return g_switches;
}

int env_int(const char* name, int fallback) { return env_raw(name, fallback); }

static DeviceInfo g_devices[64];
static std::once_flag g_device_once[64];

DeviceInfo* device_info(int* err) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= 64) { *err = e ? (int)e : 1; return nullptr; }
    DeviceInfo& d = g_devices[dev];
    std::call_once(g_device_once[dev], [&] {
        d.ordinal = dev;
        cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    });
    *err = 0;
    return &d;
}

void count_launch(int kind) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (kind >= 0 && kind < 3) g_launches_kind[kind].fetch_add(1, std::memory_order_relaxed);
}

int sm_count_or(int fallback) {
    int err = 0;
    DeviceInfo* dev = device_info(&err);
    return dev && dev->sm_count > 0 ? dev->sm_count : fallback;
}

template <typename CT>
static int launch_typed(const skr_program* p, int64_t numel, cudaStream_t stream, bool aligned) {
    KProgram<CT> k;
    memset(&k, 0, sizeof(k));
    k.n_ops = p->n_ops;
    k.n_inputs = p->n_inputs;
    k.n_outputs = p->n_outputs;
    k.numel = numel;
    for (int i = 0; i < p->n_ops; ++i) {
        const skr_op& o = p->ops[i];
        KOp<CT>& d = k.ops[i];
        d.code = o.code; d.a = o.a; d.b = o.b; d.pad = 0; d.src = o.src; d.dst = o.dst;
        for (int j = 0; j < 4; ++j) d.c[j] = (CT)o.c[j];  // the one rounding of each scalar
    }
    uint32_t off = 0;
    for (int i = 0; i < p->n_inputs; ++i) {
        k.in_ptr[i] = p->inputs[i].ptr;
        k.in_dtype[i] = (uint8_t)p->inputs[i].dtype;
        k.in_off[i] = off;
        off += kTile * dtype_size_host(p->inputs[i].dtype);
    }
    for (int i = 0; i < p->n_outputs; ++i) {
        k.out_ptr[i] = p->outputs[i].ptr;
        k.out_dtype[i] = (uint8_t)p->outputs[i].dtype;
    }
    k.stage_bytes = off;
    fill_kphilox(k.philox, p->philox, p->n_philox);

    int err = 0;
    DeviceInfo* dev = device_info(&err);
    if (!dev) return fail(err, "cudaGetDevice failed");

    const int64_t n_tiles = (numel + kTile - 1) / kTile;
    const int64_t n_full = numel / kTile;

    // Pipeline depth: keep >= ~48 KB of loads in flight per SM, two CTAs per SM when a stage is small.
    int stages = 2;
    int ctas_per_sm = 2;
    size_t smem = 0;
    k.use_tma = (aligned && n_full > 0 && off > 0) ? 1u : 0u;
    if (k.use_tma) {
        const uint32_t budget2 = 100u * 1024u;  // per CTA when two CTAs share an SM
        const uint32_t budget1 = (uint32_t)dev->max_smem - 2048u;  // leave room for the static barriers
        if (2u * off <= budget2) {
            stages = (int)(budget2 / off);
            ctas_per_sm = 2;
        } else if (2u * off <= budget1) {
            stages = (int)(budget1 / off);
            ctas_per_sm = 1;
        } else {
            k.use_tma = 0;  // a single tile of all inputs does not fit twice: element-wise path
        }
        if (stages > kMaxStages) stages = kMaxStages;
        if (stages > 4 && (uint32_t)stages * off > 64u * 1024u) {
            // enough bytes in flight already; do not hoard shared memory
            stages = (int)((64u * 1024u + off - 1) / off);
            if (stages < 3) stages = 3;
        }
    }
    k.stages = stages;
    if (k.use_tma) smem = (size_t)stages * off;

    int64_t grid = (int64_t)dev->sm_count * ctas_per_sm;
    const int64_t work = k.use_tma ? (n_full > 0 ? n_full : 1) : n_tiles;
    if (grid > work) grid = work;
    if (!k.use_tma) {
        grid = n_tiles < (int64_t)dev->sm_count * 8 ? n_tiles : (int64_t)dev->sm_count * 8;
    }
    if (grid < 1) grid = 1;

    const int which = sizeof(CT) == 8 ? 1 : 0;
    if (!dev->attr_set[which].load(std::memory_order_acquire)) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, step_kernel<CT>);
        if (e != cudaSuccess) return fail((int)e, "cudaFuncGetAttributes: %s", cudaGetErrorString(e));
        e = cudaFuncSetAttribute(step_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 dev->max_smem - (int)fa.sharedSizeBytes);
        if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        dev->attr_set[which].store(true, std::memory_order_release);
    }
    step_kernel<CT><<<(unsigned)grid, kThreads, smem, stream>>>(k);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "step kernel launch: %s", cudaGetErrorString(e));
    count_launch(1);
    return 0;
}

PipeShape pick_shape(uint32_t stage_bytes, int max_smem, int max_ctas) {
    // keep roughly 64-96 KB of bulk loads in flight per SM; more CTAs per SM when a stage is small
    PipeShape sh{2, 1, true};
    const uint32_t usable = (uint32_t)max_smem - 4096u;
    if (stage_bytes == 0 || 2u * stage_bytes > usable) { sh.ok = false; return sh; }
    int ctas = max_ctas;
    while (ctas > 1 && 2u * stage_bytes * (uint32_t)ctas > usable) --ctas;
    uint32_t per_cta = usable / (uint32_t)ctas;
    int stages = (int)(per_cta / stage_bytes);
    const int want = (int)((96u * 1024u / (uint32_t)ctas + stage_bytes - 1) / stage_bytes) + 1;
    if (stages > want) stages = want;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) stages = 2;
    const Switches& sw = switches();
    const int force_stages = sw.stages.load(std::memory_order_relaxed), force_ctas = sw.ctas.load(std::memory_order_relaxed);
    sh.stages = force_stages > 0 ? force_stages : stages;
    sh.ctas_per_sm = force_ctas > 0 ? force_ctas : ctas;
    if (sh.stages < 2) sh.stages = 2;
    if (sh.stages > kMaxStages) sh.stages = kMaxStages;
    if ((uint32_t)sh.stages * stage_bytes > usable) sh.stages = (int)(usable / stage_bytes);
    return sh;
}

// ---- choosing the block-kernel instantiation of a parsed step -----------------------------------------------
// `k` carries the control fields and the dtypes of every tensor (bind_tensors / fill_dtypes); pointers play no part.

template <typename CT, int MODE, int V>
static BlockLauncher<CT> generic_launcher(int n_philox) {
    if (n_philox > 0) return &launch_block_one<CT, MODE, V, true, ShAny>;
    return &launch_block_one<CT, MODE, V, false, ShAny>;
}

static BlockLauncher<float> pinned_any(const BProgram<float>& k, int n_philox, const char** name) {
    if (switches().no_pinned.load(std::memory_order_relaxed)) return nullptr;
    // the latent storage type is that of the network output (head.y) or, for RK combinations, of the sample
    const int probe = k.head.y_in >= 0 ? k.head.y_in : k.head.x_in;
    if (probe < 0) return nullptr;
    const bool contracted = switches().arithmetic.load(std::memory_order_relaxed) == 1;
    switch (k.in_dtype[probe]) {
        case SKR_F32: return pinned_f32(k, n_philox > 0, contracted, name);
        case SKR_BF16: return pinned_bf16(k, n_philox > 0, contracted, name);
        case SKR_F16: return pinned_f16(k, n_philox > 0, contracted, name);
        default: return nullptr;
    }
}

// offsets: a draw of this launch carries Offset noise, which only the generic shape draws (BlkAny::offsets)
static BlockLauncher<double> select_launcher(const BProgram<double>&, int n_philox, const char** name, bool = false) {
    *name = "any";
    return generic_launcher<double, IN_MIXED, 4>(n_philox);
}

static BlockLauncher<float> select_launcher(const BProgram<float>& k, int n_philox, const char** name, bool offsets = false) {
    *name = "any";
    const StorageClass storage(k);
    const int force = switches().in_mode.load(std::memory_order_relaxed);  // development switch: 0 / 8 force a mixed instantiation
    if (force == 0) return generic_launcher<float, IN_MIXED, 4>(n_philox);
    if (force == 8) return generic_launcher<float, IN_MIXED, 8>(n_philox);
    if (!offsets) {
        if (BlockLauncher<float> pinned = pinned_any(k, n_philox, name)) return pinned;
    }
    if (storage.all_f32) return generic_launcher<float, IN_F32, 4>(n_philox);
    if (storage.all_bf16) return generic_launcher<float, IN_BF16, 8>(n_philox);
    if (storage.all_f16) return generic_launcher<float, IN_F16, 8>(n_philox);
    // Mixed storage stays at 4 elements per thread: measured on B200 the 8-wide variant loses more to halved
    // occupancy / doubled stage size than it gains from amortised control (Adams-9 bf16 59 vs 31 us,
    // UniPC-3 bf16 46 vs 37 us per step); SKR_IN_MODE=8 keeps it reachable for experiments.
    return generic_launcher<float, IN_MIXED, 4>(n_philox);
}

// ---- validation shared by skr_program_launch and skr_plan_create ----------------------------------------------
// numel < 0: the size is not known yet (plan creation): size-dependent checks are left to the launch.

static int validate_program(const skr_program* p, int64_t numel, bool need_pointers) {
    if (!p) return fail(SKR_E_NULL, "null program");
    if (p->n_ops < 0 || p->n_ops > SKR_MAX_OPS) return fail(SKR_E_RANGE, "n_ops %d out of range", p->n_ops);
    if (p->n_inputs < 0 || p->n_inputs > SKR_MAX_INPUTS) return fail(SKR_E_RANGE, "n_inputs %d out of range", p->n_inputs);
    if (p->n_outputs < 0 || p->n_outputs > SKR_MAX_OUTPUTS) return fail(SKR_E_RANGE, "n_outputs %d out of range", p->n_outputs);
    if (p->n_philox < 0 || p->n_philox > SKR_MAX_PHILOX) return fail(SKR_E_RANGE, "n_philox %d out of range", p->n_philox);
    for (int i = 0; i < p->n_inputs; ++i) {
        const skr_tensor& t = p->inputs[i];
        if (t.dtype < 0 || t.dtype > SKR_F16) return fail(SKR_E_DTYPE, "input %d: unknown dtype %d", i, t.dtype);
        if (need_pointers && !t.ptr && numel > 0) return fail(SKR_E_NULL, "input %d: null pointer", i);
    }
    for (int i = 0; i < p->n_outputs; ++i) {
        const skr_tensor& t = p->outputs[i];
        if (t.dtype < 0 || t.dtype > SKR_F16) return fail(SKR_E_DTYPE, "output %d: unknown dtype %d", i, t.dtype);
        if (need_pointers && !t.ptr && numel > 0) return fail(SKR_E_NULL, "output %d: null pointer", i);
    }
    for (int i = 0; i < p->n_ops; ++i) {
        const skr_op& o = p->ops[i];
        if (o.code >= SKR_OP__COUNT) return fail(SKR_E_OPCODE, "op %d: unknown code %d", i, (int)o.code);
        const bool draws = o.code == SKR_OP_FWD && (o.b & 2);
        if (draws && (o.b & 1)) return fail(SKR_E_RANGE, "op %d: noise is either a tensor or a Philox draw", i);
        if (draws ? (o.src < 0 || o.src >= p->n_philox) : o.src >= p->n_inputs)
            return fail(SKR_E_RANGE, "op %d: input index %d out of range", i, (int)o.src);
        if (o.dst >= p->n_outputs) return fail(SKR_E_RANGE, "op %d: output index %d out of range", i, (int)o.dst);
        if (o.code == SKR_OP_STORE && o.dst < 0) return fail(SKR_E_RANGE, "op %d: STORE without an output", i);
        const bool reads = o.code == SKR_OP_LOAD || o.code == SKR_OP_UNI || o.code == SKR_OP_DPM2 || o.code == SKR_OP_DPM3A ||
                           o.code == SKR_OP_DPM3B || o.code == SKR_OP_AXPBY ||
                           ((o.code == SKR_OP_ACC0 || o.code == SKR_OP_ACC) && o.a == 0) ||
                           (o.code == SKR_OP_CONV && o.b == 0) || ((o.code == SKR_OP_FWD || o.code == SKR_OP_BACK) && (o.b & 1));
        if (reads && o.src < 0) return fail(SKR_E_RANGE, "op %d (code %d) needs an input tensor", i, (int)o.code);
        if ((o.code == SKR_OP_LOAD || o.code == SKR_OP_MOV || o.code == SKR_OP_STORE || o.code == SKR_OP_FWD) && o.a > 7)
            return fail(SKR_E_RANGE, "op %d: register %d out of range", i, (int)o.a);
        if (o.code == SKR_OP_MOV && o.b > 7) return fail(SKR_E_RANGE, "op %d: register %d out of range", i, (int)o.b);
        if ((o.code == SKR_OP_ACC0 || o.code == SKR_OP_ACC) && o.a > 8) return fail(SKR_E_RANGE, "op %d: register out of range", i);
    }
    return 0;
}

static bool draws_offsets(const skr_philox* draws, int n) {
    for (int i = 0; i < n; ++i)
        if (draws[i].offset_scale != 0.0f && draws[i].offset_inner > 0) return true;
    return false;
}

static int validate_philox(const skr_philox* draws, int n, int64_t numel) {
    for (int i = 0; i < n; ++i) {
        const skr_philox& d = draws[i];
        if (d.n_items < 1 || d.n_items > SKR_MAX_PHILOX_ITEMS) return fail(SKR_E_RANGE, "philox %d: n_items %d out of range", i, d.n_items);
        if (d.item_numel < 1 || d.item_numel * d.n_items != numel) return fail(SKR_E_SHAPE, "philox %d: n_items * item_numel != numel", i);
        if (d.offset_scale != 0.0f && (!(d.offset_scale == d.offset_scale) || d.offset_inner < 1 || d.item_numel % d.offset_inner != 0))
            return fail(SKR_E_SHAPE, "philox %d: offset_inner %lld does not divide item_numel", i, (long long)d.offset_inner);
    }
    return 0;
}

static bool any_f64(const skr_program* p) {
    bool any64 = false;
    for (int i = 0; i < p->n_inputs; ++i) any64 |= p->inputs[i].dtype == SKR_F64;
    for (int i = 0; i < p->n_outputs; ++i) any64 |= p->outputs[i].dtype == SKR_F64;
    return any64;
}

template <typename CT>
static int launch_any(const skr_program* p, int64_t numel, cudaStream_t stream, bool aligned) {
    if (!switches().force_interp.load(std::memory_order_relaxed)) {
        BProgram<CT> b;
        memset(&b, 0, sizeof(b));
        if (parse_block_program<CT>(p, b)) {
            bind_tensors(p, b);
            const char* name = nullptr;
            return select_launcher(b, p->n_philox, &name, draws_offsets(p->philox, p->n_philox))(b, p->philox, p->n_philox, numel, stream, aligned);
        }
    }
    return launch_typed<CT>(p, numel, stream, aligned);
}

}  // namespace skr

// A step whose parsing and kernel selection are done: what the plan-cache hit path of the Python layer launches.
// Immutable after skr_plan_create, so any number of threads may launch one plan at the same time.
struct skr_plan {
    int32_t n_inputs, n_outputs, n_philox;
    bool f64, block;
    const char* shape_name;
    std::unique_ptr<skr::BProgram<float>> bf;
    std::unique_ptr<skr::BProgram<double>> bd;
    skr::BlockLauncher<float> lf = nullptr, lf_offsets = nullptr;  // *_offsets: the generic shape, for draws that carry Offset noise
    skr::BlockLauncher<double> ld = nullptr;
    skr_program source;  // ops + dtypes (the interpreter's input; pointers are filled per launch)
};

template <typename CT>
static int launch_planned(const skr::BProgram<CT>& parsed, skr::BlockLauncher<CT> launcher, const skr_plan* plan, const void* const* tensors,
                          int64_t numel, const skr_philox* draws, cudaStream_t stream, bool aligned) {
    skr::BProgram<CT> k = parsed;  // the launch's own copy: plans are shared between threads
    for (int i = 0; i < plan->n_inputs; ++i) k.in_ptr[i] = tensors[i];
    for (int i = 0; i < plan->n_outputs; ++i) k.out_ptr[i] = const_cast<void*>(tensors[plan->n_inputs + i]);
    return launcher(k, draws, plan->n_philox, numel, stream, aligned);
}

extern "C" {

int skr_version(void) { return SKR_VERSION; }
const char* skr_last_error(void) { return skr::g_error; }
int64_t skr_launch_count(void) { return skr::g_launches.load(std::memory_order_relaxed); }
int64_t skr_launch_count_kind(int32_t kind) {
    return (kind >= 0 && kind <= 2) ? skr::g_launches_kind[kind].load(std::memory_order_relaxed) : -1;
}
void skr_reload_env(void) {
    skr::switches();
    skr::load_switches();
}
int skr_set_arithmetic(int32_t mode) {
    if (mode != 0 && mode != 1) return skr::fail(SKR_E_RANGE, "arithmetic mode %d: 0 (exact) or 1 (contracted)", mode);
    skr::switches();
    skr::g_switches.arithmetic.store(mode, std::memory_order_relaxed);
    return 0;
}
int skr_get_arithmetic(void) { return skr::switches().arithmetic.load(std::memory_order_relaxed); }

int skr_program_classify(const skr_program* p) {
    using namespace skr;
    if (!p) return fail(SKR_E_NULL, "null program");
    if (p->n_ops < 0 || p->n_ops > SKR_MAX_OPS) return fail(SKR_E_RANGE, "n_ops %d out of range", p->n_ops);
    auto b = std::make_unique<BProgram<double>>();  // too large for the stack of small threads
    memset(b.get(), 0, sizeof(*b));
    return parse_block_program<double>(p, *b) ? 0 : 1;
}

int skr_program_describe(const skr_program* p, char* text, int32_t capacity) {
    using namespace skr;
    if (!p || !text || capacity < 1) return fail(SKR_E_NULL, "null argument");
    if (p->n_ops < 0 || p->n_ops > SKR_MAX_OPS) return fail(SKR_E_RANGE, "n_ops %d out of range", p->n_ops);
    if (p->n_inputs < 0 || p->n_inputs > SKR_MAX_INPUTS) return fail(SKR_E_RANGE, "n_inputs %d out of range", p->n_inputs);
    if (p->n_outputs < 0 || p->n_outputs > SKR_MAX_OUTPUTS) return fail(SKR_E_RANGE, "n_outputs %d out of range", p->n_outputs);
    const bool any64 = any_f64(p);
    auto b = std::make_unique<BProgram<float>>();
    memset(b.get(), 0, sizeof(*b));
    if (!parse_block_program<float>(p, *b)) {
        snprintf(text, (size_t)capacity, "interpreter");
        return 1;
    }
    fill_dtypes(p, *b);
    const char* shape_name = "any";
    if (!any64) select_launcher(*b, p->n_philox, &shape_name);
    const BHead<float>& h = b->head;
    int n = snprintf(text, (size_t)capacity, "block compute=%s shape=%s fast_div=%d head[x=%d y=%d neg=%d conv=%d sp=%d sp2=%d]",
                     any64 ? "f64" : "f32", any64 ? "any" : shape_name, any64 ? 0 : b->fast_div,
                     h.x_in >= 0, h.y_in >= 0, h.neg, h.n_conv, h.store_p >= 0, h.store_p2 >= 0);
    for (int i = 0; i < 2 && n > 0 && n < capacity; ++i) {
        const BBlock<float>& k = b->blk[i];
        if (!k.enabled) continue;
        n += snprintf(text + n, (size_t)(capacity - n),
                      " blk%d[kind=%d sample=%d base=%d p_mode=%d div=%d pred_p=%d noise=%d terms=%d store=%d link=%d slink=%d]", i,
                      k.kind, k.sample_in >= 0, k.base_in >= 0, k.p_mode, k.has_div, k.pred_is_p, k.has_noise, k.n_terms,
                      k.store_r >= 0, k.link, k.store_link >= 0);
    }
    return 0;
}

int skr_program_launch(const skr_program* p, int64_t numel, void* stream) {
    using namespace skr;
    if (!p) return fail(SKR_E_NULL, "null program");
    if (numel < 0) return fail(SKR_E_RANGE, "negative numel");
    if (int rc = validate_program(p, numel, true)) return rc;
    if (int rc = validate_philox(p->philox, p->n_philox, numel)) return rc;
    bool aligned = true;
    for (int i = 0; i < p->n_inputs; ++i) aligned &= (reinterpret_cast<uintptr_t>(p->inputs[i].ptr) & 15u) == 0;
    for (int i = 0; i < p->n_outputs; ++i) aligned &= (reinterpret_cast<uintptr_t>(p->outputs[i].ptr) & 15u) == 0;
    if (numel == 0 || p->n_ops == 0) return 0;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    return any_f64(p) ? launch_any<double>(p, numel, s, aligned) : launch_any<float>(p, numel, s, aligned);
}

int skr_plan_create(const skr_program* p, skr_plan** out) {
    using namespace skr;
    if (!out) return fail(SKR_E_NULL, "null plan slot");
    *out = nullptr;
    if (int rc = validate_program(p, -1, false)) return rc;
    std::unique_ptr<skr_plan> plan(new (std::nothrow) skr_plan());
    if (!plan) return fail(SKR_E_UNSUPPORTED, "out of host memory");
    plan->n_inputs = p->n_inputs;
    plan->n_outputs = p->n_outputs;
    plan->n_philox = p->n_philox;
    plan->f64 = any_f64(p);
    plan->block = false;
    plan->shape_name = "interpreter";
    plan->source = *p;
    for (int i = 0; i < SKR_MAX_INPUTS; ++i) plan->source.inputs[i].ptr = nullptr;
    for (int i = 0; i < SKR_MAX_OUTPUTS; ++i) plan->source.outputs[i].ptr = nullptr;
    if (!switches().force_interp.load(std::memory_order_relaxed) && p->n_ops > 0) {
        if (plan->f64) {
            plan->bd.reset(new (std::nothrow) BProgram<double>());
            if (!plan->bd) return fail(SKR_E_UNSUPPORTED, "out of host memory");
            memset(plan->bd.get(), 0, sizeof(BProgram<double>));
            if (parse_block_program<double>(p, *plan->bd)) {
                fill_dtypes(p, *plan->bd);
                plan->ld = select_launcher(*plan->bd, p->n_philox, &plan->shape_name);
                plan->block = true;
            }
        } else {
            plan->bf.reset(new (std::nothrow) BProgram<float>());
            if (!plan->bf) return fail(SKR_E_UNSUPPORTED, "out of host memory");
            memset(plan->bf.get(), 0, sizeof(BProgram<float>));
            if (parse_block_program<float>(p, *plan->bf)) {
                fill_dtypes(p, *plan->bf);
                plan->lf = select_launcher(*plan->bf, p->n_philox, &plan->shape_name);
                const char* unused = nullptr;
                plan->lf_offsets = p->n_philox > 0 ? select_launcher(*plan->bf, p->n_philox, &unused, true) : plan->lf;
                plan->block = true;
            }
        }
    }
    *out = plan.release();
    return 0;
}

void skr_plan_destroy(skr_plan* plan) { delete plan; }

int skr_plan_kind(const skr_plan* plan) { return !plan ? skr::fail(SKR_E_NULL, "null plan") : plan->block ? 0 : 1; }

const char* skr_plan_shape(const skr_plan* plan) { return plan ? plan->shape_name : ""; }

int skr_plan_launch(const skr_plan* plan, const void* const* tensors, int64_t numel, const skr_philox* draws, void* stream) {
    using namespace skr;
    if (!plan) return fail(SKR_E_NULL, "null plan");
    if (numel < 0) return fail(SKR_E_RANGE, "negative numel");
    const int n = plan->n_inputs + plan->n_outputs;
    if (n > 0 && !tensors) return fail(SKR_E_NULL, "null tensor table");
    if (plan->n_philox > 0) {
        if (!draws) return fail(SKR_E_NULL, "the plan draws noise in the kernel: Philox keys are required");
        if (int rc = validate_philox(draws, plan->n_philox, numel)) return rc;
    }
    if (numel == 0 || plan->source.n_ops == 0) return 0;
    bool aligned = true;
    for (int i = 0; i < n; ++i) {
        if (!tensors[i]) return fail(SKR_E_NULL, "tensor %d: null pointer", i);
        aligned &= (reinterpret_cast<uintptr_t>(tensors[i]) & 15u) == 0;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (plan->block) {
        if (plan->f64) return launch_planned<double>(*plan->bd, plan->ld, plan, tensors, numel, draws, s, aligned);
        const bool offsets = plan->n_philox > 0 && draws_offsets(draws, plan->n_philox);
        return launch_planned<float>(*plan->bf, offsets ? plan->lf_offsets : plan->lf, plan, tensors, numel, draws, s, aligned);
    }
    skr_program p = plan->source;
    for (int i = 0; i < plan->n_inputs; ++i) p.inputs[i].ptr = const_cast<void*>(tensors[i]);
    for (int i = 0; i < plan->n_outputs; ++i) p.outputs[i].ptr = const_cast<void*>(tensors[plan->n_inputs + i]);
    for (int i = 0; i < plan->n_philox; ++i) p.philox[i] = draws[i];
    return plan->f64 ? launch_typed<double>(&p, numel, s, aligned) : launch_typed<float>(&p, numel, s, aligned);
}

int skr_axpby(const void* sample, const void* noise, void* out, int32_t dtype, int64_t numel, double sigma, double alpha,
              int32_t remove, void* stream) {
    skr_program p;
    memset(&p, 0, sizeof(p));
    p.n_ops = 3; p.n_inputs = 2; p.n_outputs = 1;
    p.inputs[0].ptr = const_cast<void*>(sample); p.inputs[0].dtype = dtype;
    p.inputs[1].ptr = const_cast<void*>(noise); p.inputs[1].dtype = dtype;
    p.outputs[0].ptr = out; p.outputs[0].dtype = dtype;
    p.ops[0].code = SKR_OP_LOAD; p.ops[0].a = SKR_X; p.ops[0].src = 0; p.ops[0].dst = -1;
    p.ops[1].code = SKR_OP_AXPBY; p.ops[1].a = remove ? 1 : 0; p.ops[1].src = 1; p.ops[1].dst = -1;
    p.ops[1].c[0] = remove ? sigma : alpha; p.ops[1].c[1] = remove ? alpha : sigma;
    p.ops[2].code = SKR_OP_STORE; p.ops[2].a = SKR_R; p.ops[2].src = -1; p.ops[2].dst = 0;
    return skr_program_launch(&p, numel, stream);
}

}  // extern "C"
