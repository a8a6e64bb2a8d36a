// skrample_b200 - host launcher of one block-kernel instantiation, shared by step_kernel.cu (generic shapes)
// and pinned_shapes.cu (one translation unit per latent storage type, so the instantiations compile in parallel).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstring>

#include "../../include/skrample_b200.h"
#include "block_kernel.cuh"
#include "host.cuh"

namespace skr {

// A launcher runs one block-kernel instantiation on a descriptor whose control fields, tensor tables (n_inputs,
// in_ptr / in_dtype, out_ptr / out_dtype) and Philox keys are filled in; it derives everything that depends on the
// size and the pointers (tile offsets, pipeline shape, grid) and launches.  Choosing the launcher is the expensive
// part of a launch (matching the pinned shapes): skr_plan_create does it once, skr_plan_launch only calls it.
template <typename CT>
using BlockLauncher = int (*)(BProgram<CT>& k, const skr_philox* draws, int n_draws, int64_t numel, cudaStream_t stream, bool aligned);

template <typename CT, int MODE, int V, bool PHILOX, typename Sh>
static int launch_block_one(BProgram<CT>& k, const skr_philox* draws, int n_draws, int64_t numel, cudaStream_t stream, bool aligned) {
    constexpr int TILE = kThreads * V;
    k.numel = numel;
    uint32_t off = 0;
    for (int i = 0; i < k.n_inputs; ++i) {
        k.in_off[i] = off;
        off += TILE * dtype_size_host(k.in_dtype[i]);
    }
    k.stage_bytes = off;
    resolve_offsets(k);

    int err = 0;
    DeviceInfo* dev = device_info(&err);
    if (!dev) return fail(err, "cudaGetDevice failed");

    const int64_t n_tiles = (numel + TILE - 1) / TILE;
    const int64_t n_full = numel / TILE;
    if (n_full > 0x7fffffff) return fail(SKR_E_RANGE, "numel too large");
    PipeShape sh = pick_shape(off, dev->max_smem, block_ctas_per_sm<CT, V, Sh>());
    k.use_tma = (aligned && n_tiles > 0 && sh.ok) ? 1u : 0u;
    k.n_full_tiles = (int32_t)n_full;
    k.tail_elems = (int32_t)(numel - n_full * TILE);
    k.vec_ok = aligned ? 1 : 0;
    k.stages = sh.stages;
    size_t smem = k.use_tma ? (size_t)sh.stages * off : 0;

    int64_t grid;
    if (k.use_tma) {
        grid = (int64_t)dev->sm_count * sh.ctas_per_sm;
        if (grid > n_tiles) grid = n_tiles;  // the ragged last tile is staged like the others
    } else {
        grid = n_tiles < (int64_t)dev->sm_count * 8 ? n_tiles : (int64_t)dev->sm_count * 8;
    }
    if (grid < 1) grid = 1;

    // per instantiation, one bit per device ordinal; setting the attribute twice from two threads is harmless
    static std::atomic<uint64_t> attr_set{0};
    const uint64_t bit = 1ull << (dev->ordinal & 63);
    if (!(attr_set.load(std::memory_order_acquire) & bit)) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, block_kernel<CT, MODE, V, PHILOX, Sh>);
        if (e != cudaSuccess) return fail((int)e, "cudaFuncGetAttributes: %s", cudaGetErrorString(e));
        e = cudaFuncSetAttribute(block_kernel<CT, MODE, V, PHILOX, Sh>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 dev->max_smem - (int)fa.sharedSizeBytes);
        if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set.fetch_or(bit, std::memory_order_release);
    }
    // Programmatic dependent launch: consecutive steps overlap launch latency and prologue (SKR_PDL=0 disables).
    static const bool pdl = env_int("SKR_PDL", 1) != 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads + kProducerThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    PhiloxKeys<PHILOX> keys;
    if constexpr (PHILOX) {
        memset(&keys, 0, sizeof(keys));
        fill_kphilox(keys.table, draws, n_draws < SKR_MAX_PHILOX ? n_draws : SKR_MAX_PHILOX);
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, block_kernel<CT, MODE, V, PHILOX, Sh>, k, keys);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "block kernel launch: %s", cudaGetErrorString(e));
    count_launch(0);
    return 0;
}

// Tensor tables of a launch, from the C ABI's program (pointers + dtypes).
template <typename CT>
static void bind_tensors(const skr_program* p, BProgram<CT>& k) {
    k.n_inputs = p->n_inputs;
    for (int i = 0; i < p->n_inputs; ++i) {
        k.in_ptr[i] = p->inputs[i].ptr;
        k.in_dtype[i] = p->inputs[i].dtype;
    }
    for (int i = 0; i < p->n_outputs; ++i) {
        k.out_ptr[i] = p->outputs[i].ptr;
        k.out_dtype[i] = p->outputs[i].dtype;
    }
}

template <typename Sh, int MODE, int V>
struct ShapeEntry {
    using shape = Sh;
    static constexpr int mode = MODE, v = V;
    const char* name;
};

struct StorageClass {
    bool all_f32 = true, all_bf16 = true, all_f16 = true;
    template <typename CT>
    explicit StorageClass(const BProgram<CT>& k) {
        for (int i = 0; i < k.n_inputs; ++i) {
            all_f32 &= k.in_dtype[i] == SKR_F32;
            all_bf16 &= k.in_dtype[i] == SKR_BF16;
            all_f16 &= k.in_dtype[i] == SKR_F16;
        }
    }
    bool allows(int mode) const {
        return mode == IN_MIXED || (mode == IN_F32 && all_f32) || (mode == IN_BF16 && all_bf16) || (mode == IN_F16 && all_f16);
    }
};

template <typename CT>
static void fill_dtypes(const skr_program* p, BProgram<CT>& k) {
    k.n_inputs = p->n_inputs;
    for (int i = 0; i < p->n_inputs; ++i) k.in_dtype[i] = p->inputs[i].dtype;
    for (int i = 0; i < p->n_outputs; ++i) k.out_dtype[i] = p->outputs[i].dtype;
}


// Pinned shapes of one latent storage type (pinned_shapes.cu): the launcher of the first shape that matches the
// descriptor's control fields and dtypes (its name in *name), or nullptr.
// `philox`: the step draws noise inside the kernel (instantiations with the Philox code); `contracted`: the caller
// opted into contracted arithmetic (skr_set_arithmetic) - shapes that have such an instantiation use it.
BlockLauncher<float> pinned_f32(const BProgram<float>& k, bool philox, bool contracted, const char** name);
BlockLauncher<float> pinned_bf16(const BProgram<float>& k, bool philox, bool contracted, const char** name);
BlockLauncher<float> pinned_f16(const BProgram<float>& k, bool philox, bool contracted, const char** name);

}  // namespace skr
