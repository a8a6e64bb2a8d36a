// skrample_b200 - block-kernel instantiations for the steady-state steps of the standard samplers (pinned shapes,
// block_kernel.cuh).  Compiled once per latent storage type (-DSKR_LP=0 fp32 / 2 bf16 / 3 fp16) so the three sets
// build in parallel; step_kernel.cu dispatches to pinned_f32 / pinned_bf16 / pinned_f16 before its generic kernels.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "../../include/skrample_b200.h"
#include "common.cuh"
#include "machine.cuh"
#include "block_kernel.cuh"
#include "block_launch.cuh"

#ifndef SKR_LP
#define SKR_LP 0
#endif

namespace skr {

#if SKR_LP == 0
#define SKR_PINNED_ENTRY pinned_f32
#define SKR_LP_NAME "f32"
constexpr int kLP = SKR_F32, kModeLP = IN_MIXED, kVecLP = 4;
#elif SKR_LP == 2
#define SKR_PINNED_ENTRY pinned_bf16
#define SKR_LP_NAME "bf16"
constexpr int kLP = SKR_BF16, kModeLP = IN_BF16, kVecLP = 8;
#elif SKR_LP == 3
#define SKR_PINNED_ENTRY pinned_f16
#define SKR_LP_NAME "f16"
constexpr int kLP = SKR_F16, kModeLP = IN_F16, kVecLP = 8;
#else
#error "SKR_LP must be 0 (fp32), 2 (bf16) or 3 (fp16)"
#endif

// f(entry) returns true when it handled the program; stops at the first that does.  Shapes that read fp32 solver
// state next to 16-bit latents use mixed staging at 4 elements per thread; shapes whose inputs are all of the
// latent type use 8 elements per thread for 16-bit storage (every shared-memory read 128-bit).
template <typename F>
static bool for_each_pinned_shape(F&& f) {
    return f(ShapeEntry<ShUniPC<kLP, 2>, IN_MIXED, 4>{"unipc3/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShUniPC<kLP, 1>, IN_MIXED, 4>{"unipc2/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShUniPC<kLP>, IN_MIXED, 4>{"unipc/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShUniP<kLP>, IN_MIXED, 4>{"unip/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShAcc<kLP>, IN_MIXED, 4>{"acc/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShDpm2<kLP>, IN_MIXED, 4>{"dpm2/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShDpm3<kLP>, IN_MIXED, 4>{"dpm3/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShSPC<kLP>, IN_MIXED, 4>{"spc/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShRK<kLP>, IN_MIXED, 4>{"rk/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShEuler<kLP>, kModeLP, kVecLP>{"euler/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShRKStage<kLP>, kModeLP, kVecLP>{"rk-stage/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShRKFinal<kLP>, kModeLP, kVecLP>{"rk-final/" SKR_LP_NAME});
}

// The samplers that take noise, again with the in-kernel Philox draw compiled in (the draw replaces the noise
// tensor's read here and its write by the fill kernel; for UniPC the corrector re-draws the previous step's noise).
template <typename F>
static bool for_each_pinned_philox_shape(F&& f) {
    return f(ShapeEntry<ShUniPC<kLP, 2>, IN_MIXED, 4>{"unipc3+philox/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShUniPC<kLP, 1>, IN_MIXED, 4>{"unipc2+philox/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShUniPC<kLP>, IN_MIXED, 4>{"unipc+philox/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShUniP<kLP>, IN_MIXED, 4>{"unip+philox/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShAcc<kLP>, IN_MIXED, 4>{"acc+philox/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShDpm2<kLP>, IN_MIXED, 4>{"dpm2+philox/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShDpm3<kLP>, IN_MIXED, 4>{"dpm3+philox/" SKR_LP_NAME}) ||
           f(ShapeEntry<ShEuler<kLP>, kModeLP, kVecLP>{"euler+philox/" SKR_LP_NAME});
}

// The issue-bound steps again with contracted arithmetic (machine.cuh, Policy): opt-in through skr_set_arithmetic.
template <typename F>
static bool for_each_pinned_contracted_shape(F&& f) {
    return f(ShapeEntry<Contracted<ShUniPC<kLP, 2>>, IN_MIXED, 4>{"unipc3~contracted/" SKR_LP_NAME}) ||
           f(ShapeEntry<Contracted<ShUniPC<kLP, 1>>, IN_MIXED, 4>{"unipc2~contracted/" SKR_LP_NAME}) ||
           f(ShapeEntry<Contracted<ShUniPC<kLP>>, IN_MIXED, 4>{"unipc~contracted/" SKR_LP_NAME}) ||
           f(ShapeEntry<Contracted<ShUniP<kLP>>, IN_MIXED, 4>{"unip~contracted/" SKR_LP_NAME}) ||
           f(ShapeEntry<Contracted<ShAcc<kLP>>, IN_MIXED, 4>{"acc~contracted/" SKR_LP_NAME}) ||
           f(ShapeEntry<Contracted<ShSPC<kLP>>, IN_MIXED, 4>{"spc~contracted/" SKR_LP_NAME});
}

BlockLauncher<float> SKR_PINNED_ENTRY(const BProgram<float>& k, bool philox, bool contracted, const char** name) {
    const StorageClass storage(k);
    BlockLauncher<float> found = nullptr;
    if (contracted && !philox) {
        for_each_pinned_contracted_shape([&](auto entry) {
            using E = decltype(entry);
            if (!storage.allows(E::mode) || !shape_matches<typename E::shape>(k)) return false;
            *name = entry.name;
            found = &launch_block_one<float, E::mode, E::v, false, typename E::shape>;
            return true;
        });
        if (found) return found;  // other steps keep the exact kernels
    }
    if (philox) {
        for_each_pinned_philox_shape([&](auto entry) {
            using E = decltype(entry);
            if (!storage.allows(E::mode) || !shape_matches<typename E::shape>(k)) return false;
            *name = entry.name;
            found = &launch_block_one<float, E::mode, E::v, true, typename E::shape>;
            return true;
        });
        return found;
    }
    for_each_pinned_shape([&](auto entry) {
        using E = decltype(entry);
        if (!storage.allows(E::mode) || !shape_matches<typename E::shape>(k)) return false;
        *name = entry.name;
        found = &launch_block_one<float, E::mode, E::v, false, typename E::shape>;
        return true;
    });
    return found;
}

}  // namespace skr
