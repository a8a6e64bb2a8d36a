// skrample_b200 - the structured ("block") step kernel: the fast path of skr_program_launch.
//
// Every skrample solver step has the same skeleton
//
//     head      X = sample, P = convert(X, network output)            [store P]
//     block 0   (corrector of the previous step, optional)
//     block 1   (predictor of this step)
//
// where a block is "combine P with a run of history tensors, then R = X*G + pred*D + noise*Z",
// optionally followed by a link (X = R, SPC blend, or the RK backward stage).  This kernel runs
// that skeleton with *static* control flow: the only loops are the history-term loops, whose
// operands are read from the TMA-staged shared-memory tile, so there is no per-op dispatch and the
// six state vectors live in fixed registers.  A dedicated producer warp issues the bulk copies
// (one lane per input tensor) and runs ahead of the eight consumer warps through a full/empty
// mbarrier ring; consumers never synchronise with each other.
//
// The noise term can be drawn inside the kernel (PHILOX instantiations: Philox4x32-10 + Box-Muller per element,
// bit-identical to skr_noise_fill); programs that read their noise from a tensor use instantiations without
// that code, which measurably matters for the instruction footprint of the hot loop.
//
// Instantiations: storage mode of the inputs (all fp32 / all bf16 / all fp16 / mixed, chosen per
// launch) x elements per thread (4, or 8 for 16-bit storage so every shared-memory read is
// 128-bit) x compute type (fp32, fp64) x shape.  The descriptor is plain int32/float fields in the
// kernel-parameter constant bank so every test is a uniform-datapath compare; a *pinned shape*
// (see "compile-time shapes" below, instantiated in pinned_shapes.cu) turns the fields that only
// steer control flow into compile-time constants for the steady-state step of a standard sampler.
//
// Divisions are by grid-uniform scalars and use a host-side reciprocal with a residual correction
// (machine.cuh, div_uniform: exactly the IEEE quotient, exhaustively verified).  Launches carry the
// programmatic-dependent-launch attribute: the prologue of step n+1 overlaps the tail of step n.
//
// Arithmetic is identical to the interpreter (machine.cuh): individually rounded ops in the
// reference's order.  A program that does not fit the skeleton is executed by the interpreter.
#pragma once

#include "machine.cuh"

namespace skr {

constexpr int kMaxTerms = 36;
constexpr int kProducerThreads = 32;

enum BlockKind : int32_t { BK_NONE = 0, BK_ACC = 1, BK_UNI = 2, BK_DPM2 = 3, BK_DPM3 = 4 };
enum BlockLink : int32_t { BL_NONE = 0, BL_X_FROM_R = 1, BL_BLEND = 2, BL_BACK = 3, BL_BLEND_POW = 4 };
enum InMode : int { IN_MIXED = 0, IN_F32 = 1, IN_BF16 = 2, IN_F16 = 3 };

template <typename CT>
struct BTerm {  // 16 bytes for fp32 compute: one 128-bit constant load per history term
    CT c0, c1;
    CT r0;         // RN(1/c0) where c0 is a divisor (UNI terms)
    uint32_t off;  // byte offset of the term's tensor inside a staged tile (filled at launch)
};

template <typename CT>
struct BBlock {
    int32_t enabled, kind, save_s, p_mode;  // p_mode: ACC 1 = P first, 2 = P last; UNI 1 = UniC term
    int32_t has_div, pred_is_p, has_noise, link;
    int32_t n_terms, empty_sum;
    int32_t sample_in, base_in, noise_in, store_r, store_link;  // -1 = not used
    int32_t pad;
    CT p_coef, div, gamma, delta, zeta, l0, l1, e0, e1, e2;
    CT div_r, l1_r;  // reciprocals of the divisors `div` and `l1`
    CT lp, lq;       // BL_BLEND_POW: the power and its inverse (SPC power mean, structured.py:568-575)
    uint32_t sample_off, base_off, noise_off, reserved;  // staged byte offsets of sample_in / base_in / noise_in
    BTerm<CT> terms[kMaxTerms];
    int32_t term_in[kMaxTerms];  // input index of each term (guarded path, host matching)
};

template <typename CT>
struct BHead {
    int32_t x_in, y_in, store_p, n_conv, neg;
    int32_t conv_flags[2];
    int32_t store_p2;  // optional second copy of P (e.g. fp32 solver state + 16-bit copy for the caller)
    CT conv_c[2][3];
    CT conv_r[2];  // reciprocals of the conversion divisors conv_c[.][2]
    uint32_t x_off, y_off;  // staged byte offsets of x_in / y_in
};

template <typename CT>
struct BProgram {
    int64_t numel;
    int32_t n_inputs, stages;
    uint32_t stage_bytes, use_tma;
    int32_t n_full_tiles;
    int32_t fast_div;  // every divisor has a host-side reciprocal (machine.cuh, div_uniform)
    int32_t vec_ok;    // every tensor base is 16-byte aligned: whole threads of a guarded tile may use vector accesses
    int32_t tail_elems;  // elements of the partial last tile (0: the size is a multiple of the tile)
    const void* in_ptr[SKR_MAX_INPUTS];
    void* out_ptr[SKR_MAX_OUTPUTS];
    uint32_t in_off[SKR_MAX_INPUTS];
    int32_t in_dtype[SKR_MAX_INPUTS];
    int32_t out_dtype[SKR_MAX_OUTPUTS];
    BHead<CT> head;
    BBlock<CT> blk[2];
};

// ---- staged operand fetch, V elements per thread ---------------------------------------------------

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// Shared-memory reads by 32-bit shared address: the generic-pointer window arithmetic stays out of the hot loop.
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 q;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(addr));
    return q;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 q;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(q.x), "=r"(q.y) : "r"(addr));
    return q;
}

template <typename CT, int N>
__device__ __forceinline__ void unpack_f32(uint32_t addr, CT* v) {  // N floats, 16 bytes per read
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
        const uint4 q = lds128(addr + 16 * i);
        v[4 * i] = (CT)__uint_as_float(q.x); v[4 * i + 1] = (CT)__uint_as_float(q.y);
        v[4 * i + 2] = (CT)__uint_as_float(q.z); v[4 * i + 3] = (CT)__uint_as_float(q.w);
    }
}
template <typename CT, bool BF16>
__device__ __forceinline__ void unpack_half2(uint32_t w, CT& lo, CT& hi) {
    if constexpr (BF16) {
        lo = (CT)bf16_lo(w);
        hi = (CT)bf16_hi(w);
    } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
        lo = (CT)f.x;
        hi = (CT)f.y;
    }
}
template <typename CT, int N, bool BF16>
__device__ __forceinline__ void unpack_16(uint32_t addr, CT* v) {  // N 16-bit values
    if constexpr (N == 8) {
        const uint4 q = lds128(addr);
        unpack_half2<CT, BF16>(q.x, v[0], v[1]); unpack_half2<CT, BF16>(q.y, v[2], v[3]);
        unpack_half2<CT, BF16>(q.z, v[4], v[5]); unpack_half2<CT, BF16>(q.w, v[6], v[7]);
    } else {
        const uint2 q = lds64(addr);
        unpack_half2<CT, BF16>(q.x, v[0], v[1]); unpack_half2<CT, BF16>(q.y, v[2], v[3]);
    }
}
template <typename CT, int N>
__device__ __forceinline__ void unpack_f64(uint32_t addr, CT* v) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const uint4 q = lds128(addr + 16 * i);
        v[2 * i] = (CT)__hiloint2double((int)q.y, (int)q.x);
        v[2 * i + 1] = (CT)__hiloint2double((int)q.w, (int)q.z);
    }
}

// This thread's V consecutive elements of one staged input.  `stage` is the shared address of the stage, `off`
// the input's byte offset inside it, `first_elem` = tid * V.
template <typename CT, int MODE, int V>
__device__ __forceinline__ void fetch_tile(uint32_t stage, uint32_t off, int dtype, uint32_t first_elem, CT (&v)[V]) {
    const uint32_t base = stage + off;
    if constexpr (MODE == IN_F32) {
        unpack_f32<CT, V>(base + first_elem * 4u, v);
    } else if constexpr (MODE == IN_BF16) {
        unpack_16<CT, V, true>(base + first_elem * 2u, v);
    } else if constexpr (MODE == IN_F16) {
        unpack_16<CT, V, false>(base + first_elem * 2u, v);
    } else {
        switch (dtype) {
            case SKR_F32: unpack_f32<CT, V>(base + first_elem * 4u, v); break;
            case SKR_BF16: unpack_16<CT, V, true>(base + first_elem * 2u, v); break;
            case SKR_F16: unpack_16<CT, V, false>(base + first_elem * 2u, v); break;
            default:  // SKR_F64 - only reachable in the fp64-compute instantiation
                if constexpr (sizeof(CT) == 8) unpack_f64<CT, V>(base + first_elem * 8u, v);
                break;
        }
    }
}

template <typename CT, int V>
__device__ __forceinline__ void store_tile(void* ptr, int dtype, int64_t first, const CT (&v)[V]) {
    switch (dtype) {
        case SKR_F32: {
            float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(ptr) + first);
#pragma unroll
            for (int i = 0; i < V / 4; ++i) p[i] = make_float4((float)v[4 * i], (float)v[4 * i + 1], (float)v[4 * i + 2], (float)v[4 * i + 3]);
        } break;
        case SKR_BF16: {
            uint32_t w[V / 2];
#pragma unroll
            for (int i = 0; i < V / 2; ++i) w[i] = pack_bf16((float)v[2 * i], (float)v[2 * i + 1]);
            if constexpr (V == 8) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ptr) + first) = make_uint4(w[0], w[1], w[2], w[3]);
            else *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(ptr) + first) = make_uint2(w[0], w[1]);
        } break;
        case SKR_F16: {
            uint32_t w[V / 2];
#pragma unroll
            for (int i = 0; i < V / 2; ++i) w[i] = pack_f16((float)v[2 * i], (float)v[2 * i + 1]);
            if constexpr (V == 8) *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(ptr) + first) = make_uint4(w[0], w[1], w[2], w[3]);
            else *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ptr) + first) = make_uint2(w[0], w[1]);
        } break;
        default: {  // SKR_F64 - only reachable in the fp64-compute instantiation
            if constexpr (sizeof(CT) == 8) {
                double2* p = reinterpret_cast<double2*>(reinterpret_cast<double*>(ptr) + first);
#pragma unroll
                for (int i = 0; i < V / 2; ++i) p[i] = make_double2((double)v[2 * i], (double)v[2 * i + 1]);
            }
        } break;
    }
}

// Guarded element-wise variants for the ragged tail / unaligned tensors.
template <typename CT, int V>
__device__ __forceinline__ void fetch_guarded(const void* ptr, int dtype, int64_t first, int64_t numel, CT (&v)[V]) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
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

template <typename CT, int V>
__device__ __forceinline__ void store_guarded(void* ptr, int dtype, int64_t first, int64_t numel, const CT (&v)[V]) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
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

// ---- compile-time shapes -------------------------------------------------------------------------------
//
// The descriptor fields that only steer control flow (which operands exist, block kind, storage types) can be
// pinned at compile time: a shape type carries them as constants, -1 meaning "read the descriptor".  ShAny pins
// nothing and runs every program the parser accepts; the pinned shapes cover the steady-state step of the common
// samplers, where the uniform branches and dtype switches otherwise make up ~40% of the instruction stream.

template <int PIN>
__device__ __forceinline__ int pinned(int runtime) {
    if constexpr (PIN >= 0) return PIN;
    else return runtime;
}

struct BlkAny {
    static constexpr int enabled = -1, kind = -1, sample = -1, base = -1, p_mode = -1, has_div = -1, pred_p = -1;
    static constexpr int noise = -1, store = -1, link = -1, slink = -1;
    static constexpr int n_terms = -1;  // >= 0: the history loop is unrolled, its constants become immediates
    static constexpr bool offsets = true;  // an in-kernel draw may carry Offset noise (machine.cuh: draw_normals); the
                                           // pinned blocks compile it out - a launch with offsets takes the generic shape
    static constexpr int dt_state = -1, dt_sample = -1, dt_noise = -1, dt_store = -1, dt_slink = -1;
};
struct BlkOff : BlkAny {
    static constexpr int enabled = 0;
};
struct ShAny {
    static constexpr bool contract = false;  // arithmetic policy (machine.cuh): exact unless a shape opts in
    static constexpr int ctas8 = 2;  // CTAs per SM the 8-elements-per-thread instantiation is compiled for
    static constexpr int x = -1, y = -1, neg = -1, n_conv = -1, sp = -1, sp2 = -1;
    static constexpr int dt_x = -1, dt_y = -1, dt_sp = -1, dt_sp2 = -1;
    using B0 = BlkAny;
    using B1 = BlkAny;
};

// Pinned shapes.  LP is the storage type of the latents the caller hands in and gets back (sample, network
// output, noise, `final`); ST that of the solver state a block reads (x-hat history, previous samples: fp32; the
// raw derivatives of an unconverted RK step: LP).  Whether a block adds noise and whether x-hat is stored a second
// time stay run-time flags: one uniform branch each.
template <int KIND, int SAMPLE, int BASE, int PMODE, int DIV, int PREDP, int STORE, int LINK, int SLINK, int OUT, int ST, int LP, int NT = -1>
struct BlkPin : BlkAny {
    static constexpr bool offsets = false;
    static constexpr int enabled = 1, kind = KIND, sample = SAMPLE, base = BASE, p_mode = PMODE, has_div = DIV;
    static constexpr int pred_p = PREDP, store = STORE, link = LINK, slink = SLINK, n_terms = NT;
    static constexpr int dt_state = ST, dt_sample = ST, dt_noise = LP, dt_store = OUT, dt_slink = SKR_F32;
};
// explicit RK on converted derivatives (the default: derivative_transform = DataModel): X = the step's sample (LP),
// A = (sum k_i c_i [+ P c_s]) [/ sum c], derivatives k_i in fp32; covers stage inputs (has_div) and the final update
template <int LP>
struct BlkRK : BlkAny {
    static constexpr bool offsets = false;
    static constexpr int enabled = 1, kind = BK_ACC, sample = 1, base = 0, pred_p = 0, store = 1, link = BL_NONE, slink = 0;
    static constexpr int dt_state = SKR_F32, dt_sample = LP, dt_noise = LP, dt_store = LP;
};
// head of a structured sampler step: X = sample, P = x-hat = convert(X, network output), stored as fp32 state
template <int LP, typename BLK0, typename BLK1>
struct ShStep : ShAny {
    static constexpr int x = 1, y = 1, neg = 0, n_conv = 1, sp = 1;
    static constexpr int dt_x = LP, dt_y = LP, dt_sp = SKR_F32;
    using B0 = BLK0;
    using B1 = BLK1;
};
// head without a conversion: X = sample, optionally P = raw network output (Euler; RK combinations)
template <int LP, int Y, typename BLK0>
struct ShRaw : ShAny {
    static constexpr int ctas8 = 4;  // few live values: the 8-wide instantiation still fits 56 registers
    static constexpr int x = 1, y = Y, neg = 0, n_conv = 0, sp = 0, sp2 = 0;
    static constexpr int dt_x = LP, dt_y = LP;
    using B0 = BLK0;
    using B1 = BlkOff;
};

// Euler: R = X*G + y*D (+ noise*Z), the conversion folded into the scalars
template <int LP>
using ShEuler = ShRaw<LP, 1, BlkPin<BK_NONE, 0, 0, 0, 0, 1, 1, BL_NONE, 0, LP, SKR_F32, LP>>;
// Adams / DPM-1 and every "weighted sum of the x-hat history, x-hat first" predictor
template <int LP>
using ShAcc = ShStep<LP, BlkPin<BK_ACC, 0, 0, 1, 0, 0, 1, BL_NONE, 0, LP, SKR_F32, LP>, BlkOff>;
template <int LP>
using ShDpm2 = ShStep<LP, BlkPin<BK_DPM2, 0, 0, 0, 0, 0, 1, BL_NONE, 0, LP, SKR_F32, LP>, BlkOff>;
template <int LP>
using ShDpm3 = ShStep<LP, BlkPin<BK_DPM3, 0, 0, 0, 0, 0, 1, BL_NONE, 0, LP, SKR_F32, LP>, BlkOff>;
// UniP: predictor only
template <int LP>
using ShUniP = ShStep<LP, BlkPin<BK_UNI, 0, 0, 0, 0, 0, 1, BL_NONE, 0, LP, SKR_F32, LP>, BlkOff>;
// UniPC steady state: block 0 corrects the previous step (UniC term, fp32 state out, X = R), block 1 predicts.
// NT = history terms of each block (order - 1 at steady state): pinned for orders 2 and 3, a loop otherwise.
template <int LP, int NT = -1>
using ShUniPC = ShStep<LP, BlkPin<BK_UNI, 1, 1, 1, 0, 0, 1, BL_X_FROM_R, 0, SKR_F32, SKR_F32, LP, NT>,
                       BlkPin<BK_UNI, 0, 0, 0, 0, 0, 1, BL_NONE, 0, LP, SKR_F32, LP, NT>>;
// SPC steady state: block 0 = Adams corrector blended into the previous sample (fp32 state out), block 1 = Euler
template <int LP>
using ShSPC = ShStep<LP, BlkPin<BK_ACC, 1, 0, 1, 0, 0, 0, BL_BLEND, 1, SKR_F32, SKR_F32, LP>,
                     BlkPin<BK_NONE, 0, 0, 0, 0, 1, 1, BL_NONE, 0, LP, SKR_F32, LP>>;
template <int LP>
using ShRK = ShStep<LP, BlkRK<LP>, BlkOff>;
// explicit RK on raw derivatives: stage input = X*G + (sum k_i c_i / sum c_i)*D, final = X*G + (sum k_i b_i)*D
template <int LP>
using ShRKStage = ShRaw<LP, 0, BlkPin<BK_ACC, 0, 0, 0, 1, 0, 1, BL_NONE, 0, LP, LP, LP>>;
template <int LP>
using ShRKFinal = ShRaw<LP, 0, BlkPin<BK_ACC, 0, 0, 0, 0, 0, 1, BL_NONE, 0, LP, LP, LP>>;

// The same shape with contracted arithmetic (fp32 compute only).
template <typename Base>
struct Contracted : Base {
    static constexpr bool contract = true;
};

template <typename BS, typename CT>
static bool block_matches(const BBlock<CT>& k, const int32_t* in_dt, const int32_t* out_dt) {
    auto eq = [](int pin, int v) { return pin < 0 || pin == v; };
    if (!eq(BS::enabled, k.enabled)) return false;
    if (!k.enabled) return true;
    bool ok = eq(BS::kind, k.kind) && eq(BS::sample, k.sample_in >= 0) && eq(BS::base, k.base_in >= 0) &&
              eq(BS::p_mode, k.p_mode) && eq(BS::has_div, k.has_div) && eq(BS::pred_p, k.pred_is_p) &&
              eq(BS::noise, k.has_noise) && eq(BS::store, k.store_r >= 0) && eq(BS::link, k.link) &&
              eq(BS::slink, k.store_link >= 0) && eq(BS::n_terms, k.n_terms);
    if (!ok) return false;
    if (BS::dt_sample >= 0 && k.sample_in >= 0 && in_dt[k.sample_in] != BS::dt_sample) return false;
    if (BS::dt_state >= 0) {
        if (k.base_in >= 0 && in_dt[k.base_in] != BS::dt_state) return false;
        for (int t = 0; t < k.n_terms; ++t)
            if (in_dt[k.term_in[t]] != BS::dt_state) return false;
    }
    if (BS::dt_noise >= 0 && k.has_noise == 1 && in_dt[k.noise_in] != BS::dt_noise) return false;
    if (BS::dt_store >= 0 && k.store_r >= 0 && out_dt[k.store_r] != BS::dt_store) return false;
    if (BS::dt_slink >= 0 && k.store_link >= 0 && out_dt[k.store_link] != BS::dt_slink) return false;
    return true;
}

template <typename Sh, typename CT>
static bool shape_matches(const BProgram<CT>& p) {
    auto eq = [](int pin, int v) { return pin < 0 || pin == v; };
    const BHead<CT>& h = p.head;
    if (!(eq(Sh::x, h.x_in >= 0) && eq(Sh::y, h.y_in >= 0) && eq(Sh::neg, h.neg) && eq(Sh::n_conv, h.n_conv) &&
          eq(Sh::sp, h.store_p >= 0) && eq(Sh::sp2, h.store_p2 >= 0)))
        return false;
    if (Sh::dt_x >= 0 && h.x_in >= 0 && p.in_dtype[h.x_in] != Sh::dt_x) return false;
    if (Sh::dt_y >= 0 && h.y_in >= 0 && p.in_dtype[h.y_in] != Sh::dt_y) return false;
    if (Sh::dt_sp >= 0 && h.store_p >= 0 && p.out_dtype[h.store_p] != Sh::dt_sp) return false;
    if (Sh::dt_sp2 >= 0 && h.store_p2 >= 0 && p.out_dtype[h.store_p2] != Sh::dt_sp2) return false;
    return block_matches<typename Sh::B0>(p.blk[0], p.in_dtype, p.out_dtype) &&
           block_matches<typename Sh::B1>(p.blk[1], p.in_dtype, p.out_dtype);
}

// ---- one tile of the skeleton -------------------------------------------------------------------------

// Where a tile's operands come from: the staged shared-memory tile (hot loop), global memory element by element with
// bounds checks (unaligned tensors, the one thread that straddles the end), or global memory with vector accesses
// (whole threads of the ragged tail tile - independent loads the compiler can batch).
enum TilePath : int { TP_STAGED = 0, TP_ELEMENTS = 1, TP_VECTOR = 2 };

template <typename CT, int V>
__device__ __forceinline__ void fetch_vector(const void* ptr, int dtype, int64_t first, CT (&v)[V]) {
    switch (dtype) {
        case SKR_F32: {
            const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ptr) + first);
#pragma unroll
            for (int i = 0; i < V / 4; ++i) {
                const float4 q = p[i];
                v[4 * i] = (CT)q.x; v[4 * i + 1] = (CT)q.y; v[4 * i + 2] = (CT)q.z; v[4 * i + 3] = (CT)q.w;
            }
        } break;
        case SKR_BF16:
        case SKR_F16: {
            uint32_t w[V / 2];
            if constexpr (V == 8) {
                const uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(ptr) + first);
                w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
            } else {
                const uint2 q = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(ptr) + first);
                w[0] = q.x; w[1] = q.y;
            }
#pragma unroll
            for (int i = 0; i < V / 2; ++i) {
                if (dtype == SKR_BF16) unpack_half2<CT, true>(w[i], v[2 * i], v[2 * i + 1]);
                else unpack_half2<CT, false>(w[i], v[2 * i], v[2 * i + 1]);
            }
        } break;
        default: {
            if constexpr (sizeof(CT) == 8) {
                const double2* p = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(ptr) + first);
#pragma unroll
                for (int i = 0; i < V / 2; ++i) {
                    const double2 q = p[i];
                    v[2 * i] = (CT)q.x;
                    v[2 * i + 1] = (CT)q.y;
                }
            }
        } break;
    }
}

template <typename CT, int MODE, int V, int PATH, bool PARTIAL = false>
struct TileIO {
    const BProgram<CT>& prog;
    uint32_t stage;       // shared address of the staged tile
    uint32_t first_elem;  // tid * V
    int64_t first;  // PARTIAL (the ragged last tile): threads that straddle the end store element by element
    // `in` = input index (pointer / dtype tables), `off` = its byte offset inside the staged tile
    template <int DT = -1>
    __device__ __forceinline__ void load(int in, uint32_t off, CT (&v)[V]) const {
        if constexpr (PATH == TP_ELEMENTS) fetch_guarded<CT, V>(prog.in_ptr[in], prog.in_dtype[in], first, prog.numel, v);
        else if constexpr (PATH == TP_VECTOR) fetch_vector<CT, V>(prog.in_ptr[in], prog.in_dtype[in], first, v);
        else fetch_tile<CT, MODE, V>(stage, off, pinned<DT>(prog.in_dtype[in]), first_elem, v);
    }
    template <int DT = -1>
    __device__ __forceinline__ void store(int out, const CT (&v)[V]) const {
        if constexpr (PATH == TP_ELEMENTS) store_guarded<CT, V>(prog.out_ptr[out], prog.out_dtype[out], first, prog.numel, v);
        else if (PARTIAL && first + V > prog.numel) store_guarded<CT, V>(prog.out_ptr[out], prog.out_dtype[out], first, prog.numel, v);
        else store_tile<CT, V>(prog.out_ptr[out], pinned<DT>(prog.out_dtype[out]), first, v);
    }
};

template <typename CT, int V, bool CONTRACT>
__device__ __forceinline__ void divide(CT (&a)[V], CT d, CT r, bool fast) {
    if constexpr (CONTRACT && sizeof(CT) == 4) div_reciprocal<V>(a, d, r, fast);
    else div_uniform<V>(a, d, r, fast);
}

template <typename CT, int MODE, int V, int PATH, bool PHILOX, typename BS, bool PARTIAL, bool CONTRACT>
__device__ __forceinline__ void run_one_block(const BProgram<CT>& prog, const PhiloxKeys<PHILOX>& keys, const BBlock<CT>& k,
                                              const TileIO<CT, MODE, V, PATH, PARTIAL>& io, CT (&X)[V], CT (&P)[V],
                                              CT (&B)[V], CT (&A)[V], CT (&S)[V], CT (&R)[V]) {
    using Ar = Policy<CT, CONTRACT && sizeof(CT) == 4>;
    if (!pinned<BS::enabled>(k.enabled)) return;
    const bool fast_div = prog.fast_div != 0;
    const int link = pinned<BS::link>(k.link);
    if (link == BL_BLEND || link == BL_BLEND_POW) {
#pragma unroll
        for (int j = 0; j < V; ++j) S[j] = X[j];
    }
    if (pinned<BS::sample>(k.sample_in >= 0)) io.template load<BS::dt_sample>(k.sample_in, k.sample_off, X);

    CT in[V];
    const int kind = pinned<BS::kind>(k.kind);
    if (kind != BK_NONE) {
        const int n_terms = pinned<BS::n_terms>(k.n_terms);
        const int p_mode = pinned<BS::p_mode>(k.p_mode);
        if (kind != BK_ACC) {
            if (pinned<BS::base>(k.base_in >= 0)) io.template load<BS::dt_state>(k.base_in, k.base_off, B);
            else {
#pragma unroll
                for (int j = 0; j < V; ++j) B[j] = P[j];
            }
        }
        if (kind == BK_ACC) {
            int t = 0;
            if (p_mode == 1 || n_terms == 0) {
                const CT c = k.p_coef;
#pragma unroll
                for (int j = 0; j < V; ++j) A[j] = Ar::head(P[j], c);
            } else {
                io.template load<BS::dt_state>(k.term_in[0], k.terms[0].off, in);
                const CT c = k.terms[0].c0;
#pragma unroll
                for (int j = 0; j < V; ++j) A[j] = Ar::head(in[j], c);
                t = 1;
            }
            for (; t < n_terms; ++t) {
                io.template load<BS::dt_state>(k.term_in[t], k.terms[t].off, in);
                const CT c = k.terms[t].c0;
#pragma unroll
                for (int j = 0; j < V; ++j) A[j] = Ar::madd(A[j], in[j], c);
            }
            if (p_mode == 2 && n_terms > 0) {
                const CT c = k.p_coef;
#pragma unroll
                for (int j = 0; j < V; ++j) A[j] = Ar::madd(A[j], P[j], c);
            }
            if (pinned<BS::has_div>(k.has_div)) {
                divide<CT, V, CONTRACT>(A, k.div, k.div_r, fast_div);
            }
        } else if (kind == BK_UNI) {
#pragma unroll
            for (int j = 0; j < V; ++j) A[j] = (CT)0;  // 0 + first term, like the reference's running sum
            auto term = [&](int t) {
                io.template load<BS::dt_state>(k.term_in[t], k.terms[t].off, in);
                const CT rho = k.terms[t].c1;
#pragma unroll
                for (int j = 0; j < V; ++j) in[j] = Ar::sub(in[j], B[j]);
                if constexpr (Ar::contract) {
                    if (fast_div && k.terms[t].r0 != (CT)0) {  // ((x - B) / rk) * rho as one multiply-add with the constant rho / rk
                        const CT scale = Ar::mul(rho, k.terms[t].r0);
#pragma unroll
                        for (int j = 0; j < V; ++j) A[j] = Ar::madd(A[j], in[j], scale);
                        return;
                    }
                }
                divide<CT, V, CONTRACT>(in, k.terms[t].c0, k.terms[t].r0, fast_div);
#pragma unroll
                for (int j = 0; j < V; ++j) A[j] = Ar::madd(A[j], in[j], rho);
            };
            if constexpr (BS::n_terms >= 0) {
#pragma unroll
                for (int t = 0; t < BS::n_terms; ++t) term(t);
            } else {
                for (int t = 0; t < n_terms; ++t) term(t);
            }
            if (p_mode == 1) {
                const CT rho = k.p_coef;
#pragma unroll
                for (int j = 0; j < V; ++j) A[j] = Ar::madd(A[j], Ar::sub(P[j], B[j]), rho);
            }
            const bool empty = k.empty_sum != 0;
#pragma unroll
            for (int j = 0; j < V; ++j) A[j] = Ar::add(B[j], empty ? (CT)0 : A[j]);
        } else if (kind == BK_DPM2) {
            io.template load<BS::dt_state>(k.term_in[0], k.terms[0].off, in);
            const CT inv_r = k.terms[0].c0, half = k.terms[0].c1;
#pragma unroll
            for (int j = 0; j < V; ++j) A[j] = Ar::add(B[j], Ar::mul(half, Ar::mul(inv_r, Ar::sub(B[j], in[j]))));
        } else {  // BK_DPM3
            CT in2[V];
            io.template load<BS::dt_state>(k.term_in[0], k.terms[0].off, in);
            io.template load<BS::dt_state>(k.term_in[1], k.terms[1].off, in2);
            const CT inv_r = k.terms[0].c0, inv_r2 = k.terms[1].c0, mix = k.terms[1].c1;
            const CT inv_sum = k.e0, w1 = k.e1, w2 = k.e2;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const CT d10 = Ar::mul(inv_r, Ar::sub(B[j], in[j]));
                const CT d11 = Ar::mul(inv_r2, Ar::sub(in[j], in2[j]));
                const CT d = Ar::sub(d10, d11);
                const CT d1 = Ar::add(d10, Ar::mul(mix, d));
                const CT d2 = Ar::mul(inv_sum, d);
                A[j] = Ar::add(Ar::add(B[j], Ar::mul(w1, d1)), Ar::mul(w2, d2));
            }
        }
    }

    const CT gamma = k.gamma, delta = k.delta;
    const bool from_p = pinned<BS::pred_p>(k.pred_is_p) != 0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const CT pred = from_p ? P[j] : A[j];
        R[j] = Ar::madd(Ar::head(X[j], gamma), pred, delta);
    }
    const int has_noise = pinned<BS::noise>(k.has_noise);
    if (has_noise) {
        bool drawn = false;
        if constexpr (PHILOX) {
            if (has_noise == 2) {
                draw_normals<CT, V, BS::offsets>(keys.table[k.noise_in], io.first, prog.numel, in);
                drawn = true;
            }
        }
        if (!drawn) io.template load<BS::dt_noise>(k.noise_in, k.noise_off, in);
        const CT zeta = k.zeta;
#pragma unroll
        for (int j = 0; j < V; ++j) R[j] = Ar::madd(R[j], in[j], zeta);
    }
    if (pinned<BS::store>(k.store_r >= 0)) io.template store<BS::dt_store>(k.store_r, R);

    if (link != BL_NONE) {
        if (link == BL_X_FROM_R) {
#pragma unroll
            for (int j = 0; j < V; ++j) X[j] = R[j];
        } else if (link == BL_BLEND) {
            const CT l0 = k.l0, l1 = k.l1;
#pragma unroll
            for (int j = 0; j < V; ++j) X[j] = Ar::madd(Ar::mul(S[j], l0), R[j], l1);
            if (pinned<BS::slink>(k.store_link >= 0)) io.template store<BS::dt_slink>(k.store_link, X);
        } else if (link == BL_BLEND_POW) {
            // X = spow(spow(S, pw)*p + spow(R, pw)*c, 1/pw): the signed power mean, the only nonlinear op of a step
            const CT l0 = k.l0, l1 = k.l1, pw = k.lp, inv = k.lq;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const CT mixed = Ar::add(Ar::mul(Ar::spow(S[j], pw), l0), Ar::mul(Ar::spow(R[j], pw), l1));
                X[j] = Ar::spow(mixed, inv);
            }
            if (pinned<BS::slink>(k.store_link >= 0)) io.template store<BS::dt_slink>(k.store_link, X);
        } else {  // BL_BACK
            const CT l0 = k.l0;
#pragma unroll
            for (int j = 0; j < V; ++j) P[j] = Ar::msub(R[j], X[j], l0);
            divide<CT, V, CONTRACT>(P, k.l1, k.l1_r, fast_div);
            if (pinned<BS::slink>(k.store_link >= 0)) io.template store<BS::dt_slink>(k.store_link, P);
        }
    }
}

template <typename CT, int MODE, int V, int PATH, bool PHILOX, typename Sh, bool PARTIAL = false>
__device__ __forceinline__ void run_block_tile(const BProgram<CT>& prog, const PhiloxKeys<PHILOX>& keys, int64_t first, uint32_t stage, int tid) {
    constexpr bool CONTRACT = Sh::contract && sizeof(CT) == 4;
    using Ar = Policy<CT, CONTRACT>;
    const TileIO<CT, MODE, V, PATH, PARTIAL> io{prog, stage, (uint32_t)tid * V, first};

    CT X[V], P[V], B[V], A[V], S[V], R[V];
#pragma unroll
    for (int j = 0; j < V; ++j) X[j] = P[j] = B[j] = A[j] = S[j] = R[j] = (CT)0;

    // ---- head ---------------------------------------------------------------------------------
    const BHead<CT>& h = prog.head;
    const bool fast_div = prog.fast_div != 0;
    if (pinned<Sh::x>(h.x_in >= 0)) io.template load<Sh::dt_x>(h.x_in, h.x_off, X);
    if (pinned<Sh::y>(h.y_in >= 0)) {
        io.template load<Sh::dt_y>(h.y_in, h.y_off, P);
        const bool neg = pinned<Sh::neg>(h.neg) != 0;
#pragma unroll
        for (int j = 0; j < V; ++j) P[j] = neg ? -P[j] : P[j];
    }
    const int n_conv = pinned<Sh::n_conv>(h.n_conv);
    if (n_conv > 0) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (c < n_conv) {
                const int f = h.conv_flags[c];
                const CT c0 = h.conv_c[c][0], c1 = h.conv_c[c][1], c2 = h.conv_c[c][2];
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    if (f & SKR_CONV_USE_X) {
                        const CT lhs = (f & SKR_CONV_MUL_X) ? Ar::mul(c0, X[j]) : X[j];
                        P[j] = (f & SKR_CONV_MUL_Y) ? Ar::msub(lhs, c1, P[j]) : Ar::sub(lhs, P[j]);
                    } else {
                        P[j] = (f & SKR_CONV_MUL_Y) ? Ar::mul(P[j], c1) : P[j];
                    }
                }
                if (f & SKR_CONV_DIV) divide<CT, V, CONTRACT>(P, c2, h.conv_r[c], fast_div);
            }
        }
    }
    if (pinned<Sh::sp>(h.store_p >= 0)) io.template store<Sh::dt_sp>(h.store_p, P);
    if (pinned<Sh::sp2>(h.store_p2 >= 0)) io.template store<Sh::dt_sp2>(h.store_p2, P);

    run_one_block<CT, MODE, V, PATH, PHILOX, typename Sh::B0, PARTIAL, CONTRACT>(prog, keys, prog.blk[0], io, X, P, B, A, S, R);
    run_one_block<CT, MODE, V, PATH, PHILOX, typename Sh::B1, PARTIAL, CONTRACT>(prog, keys, prog.blk[1], io, X, P, B, A, S, R);
}

// Launches that cannot be staged (unaligned tensor bases, a stage too large for shared memory) run out of line so
// the pipelined loop stays compact.  Tile t belongs to CTA t % grid.  Threads whose V elements are all inside an
// aligned tensor use vector accesses; the thread straddling the end, or every thread of an unaligned launch, goes
// element by element.  One warp needs 5-20 us for a tile on this path (a long dependent instruction stream), which
// is why the ragged tail of a staged launch goes through the pipeline instead (block_kernel).
template <typename CT, int MODE, int V, bool PHILOX>
__device__ __noinline__ void run_guarded_tiles(const BProgram<CT>& prog, const PhiloxKeys<PHILOX>& keys, int64_t first_tile, int64_t n_tiles, int tid) {
    const int64_t grid = gridDim.x;
    const int64_t start = ((int64_t)blockIdx.x + grid - first_tile % grid) % grid;
    for (int64_t tile = first_tile + start; tile < n_tiles; tile += grid) {
        const int64_t first = tile * (kThreads * V) + (int64_t)tid * V;
        // This path reads operands where the step needs them, one dependent miss after the other; touching this
        // thread's slice of every input first makes the misses overlap (measured: 12.7 -> see DESIGN on a 16 MB step).
        if (first < prog.numel) {
            for (int i = 0; i < prog.n_inputs; ++i) {
                const char* at = reinterpret_cast<const char*>(prog.in_ptr[i]) + first * (int64_t)dtype_size(prog.in_dtype[i]);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(at));
            }
        }
        if (prog.vec_ok && first + V <= prog.numel) run_block_tile<CT, MODE, V, TP_VECTOR, PHILOX, ShAny>(prog, keys, first, 0u, tid);
        else if (first < prog.numel) run_block_tile<CT, MODE, V, TP_ELEMENTS, PHILOX, ShAny>(prog, keys, first, 0u, tid);
    }
}

// Occupancy each instantiation is compiled for (register cap) and launched with (pipeline shape).
template <typename CT, int V, typename Sh>
constexpr int block_ctas_per_sm() {
    return sizeof(CT) == 8 ? 2 : V == 8 ? Sh::ctas8 : 4;
}

template <typename CT, int MODE, int V, bool PHILOX, typename Sh>
__global__ void __launch_bounds__(kThreads + kProducerThreads, block_ctas_per_sm<CT, V, Sh>()) block_kernel(const __grid_constant__ BProgram<CT> prog, const __grid_constant__ PhiloxKeys<PHILOX> keys) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];

    constexpr int TILE = kThreads * V;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // staged tiles: the full ones and, when the size is ragged, the partial last one (tail_elems of it are real)
    const int n_full = prog.n_full_tiles;
    const int n_staged = prog.use_tma ? n_full + (prog.tail_elems > 0 ? 1 : 0) : 0;
    const int stages = prog.stages;
    const uint32_t stage_bytes = prog.stage_bytes;
    const int grid = (int)gridDim.x, cta = (int)blockIdx.x;
    const int mine = n_staged > cta ? (n_staged - cta + grid - 1) / grid : 0;
    // the partial tile has the highest index: it is the last tile of the CTA that owns it
    const bool own_partial = n_staged > n_full && n_full % grid == cta;
    const int mine_full = mine - (own_partial ? 1 : 0);
    const bool producer = warp == kThreads / 32;

    // PDL: the next step's grid may start its prologue now; this grid touches global memory only after every
    // predecessor has completed (the step reads what the previous step wrote).
    griddep_launch_dependents();
    if (mine > 0) {
        if (tid == 0) {
            for (int s = 0; s < stages; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], kThreads / 32);
            }
            fence_barrier_init();
        }
        __syncthreads();
    }
    griddep_wait();
    if (mine > 0) {
        if (producer) {
            // one lane per input tensor; lanes beyond n_inputs idle (SKR_MAX_INPUTS == 32 == warp size)
            const bool active = lane < prog.n_inputs;
            const uint32_t esize = active ? dtype_size(prog.in_dtype[lane]) : 0u;
            const unsigned char* src = active ? reinterpret_cast<const unsigned char*>(prog.in_ptr[lane]) : nullptr;
            unsigned char* dst = smem + (active ? prog.in_off[lane] : 0u);
            const uint32_t bytes = TILE * esize;
            src += (size_t)cta * bytes;
            const size_t stride = (size_t)grid * bytes;
            int s = 0;
            uint32_t phase = 1;  // a fresh "empty" barrier counts as already released
            for (int k = 0; k < mine_full; ++k) {
                mbar_wait(&empty_bar[s], phase);
                if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
                __syncwarp();
                if (active) tma_load_1d(dst + (size_t)s * stage_bytes, src, bytes, &full_bar[s]);
                src += stride;
                if (++s == stages) { s = 0; phase ^= 1u; }
            }
            if (own_partial) {
                // The ragged tail: a bulk copy moves multiples of 16 bytes, so each lane brings the whole 16-byte
                // chunks of its tensor's tail by TMA and the few elements after them with ordinary loads.  The
                // consumers then run the same staged code as for every other tile.
                mbar_wait(&empty_bar[s], phase);
                unsigned char* to = dst + (size_t)s * stage_bytes;
                const uint32_t real = (uint32_t)prog.tail_elems * esize;
                const uint32_t bulk = real & ~15u;
                for (uint32_t at = bulk; at < real; at += 2u)  // element sizes are multiples of 2 bytes
                    *reinterpret_cast<uint16_t*>(to + at) = *reinterpret_cast<const uint16_t*>(src + at);
                uint32_t total = bulk;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                __syncwarp();  // the stores above are ordered before lane 0's arrive (release) below
                if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], total);
                __syncwarp();
                if (active && bulk) tma_load_1d(to, src, bulk, &full_bar[s]);
            }
        } else {
            int s = 0;
            uint32_t phase = 0;
            int64_t first = (int64_t)cta * TILE + (int64_t)tid * V;
            const int64_t stride = (int64_t)grid * TILE;
            const uint32_t smem_addr = (uint32_t)__cvta_generic_to_shared(smem);
            uint32_t stage_addr = smem_addr;
            for (int k = 0; k < mine_full; ++k) {
                mbar_wait(&full_bar[s], phase);
                run_block_tile<CT, MODE, V, TP_STAGED, PHILOX, Sh>(prog, keys, first, stage_addr, tid);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
                first += stride;
                stage_addr += stage_bytes;
                if (++s == stages) { s = 0; phase ^= 1u; stage_addr = smem_addr; }
            }
            if (own_partial) {
                mbar_wait(&full_bar[s], phase);
                if (first < prog.numel) run_block_tile<CT, MODE, V, TP_STAGED, PHILOX, Sh, true>(prog, keys, first, stage_addr, tid);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
            }
        }
    }
    if (!producer && !prog.use_tma) {  // unaligned tensors or a stage that does not fit: everything element-wise
        const int64_t n_tiles = (prog.numel + TILE - 1) / TILE;
        run_guarded_tiles<CT, MODE, V, PHILOX>(prog, keys, 0, n_tiles, tid);
    }
}

// ------------------------------------------------------------------------------------------------
// host: recognise the head / block / block skeleton in a generic program

struct OpCursor {
    const skr_program* p;
    int i;
    const skr_op* peek() const { return i < p->n_ops ? &p->ops[i] : nullptr; }
    bool is(int code) const { const skr_op* o = peek(); return o && o->code == code; }
    bool is_load(int reg) const { const skr_op* o = peek(); return o && o->code == SKR_OP_LOAD && o->a == reg; }
    bool is_mov(int dst, int src) const { const skr_op* o = peek(); return o && o->code == SKR_OP_MOV && o->a == dst && o->b == src; }
    bool is_store(int reg) const { const skr_op* o = peek(); return o && o->code == SKR_OP_STORE && o->a == reg; }
    const skr_op* take() { return &p->ops[i++]; }
};

template <typename CT>
static bool parse_block(OpCursor& cur, BBlock<CT>& k) {
    memset(&k, 0, sizeof(k));
    k.sample_in = k.base_in = k.noise_in = k.store_r = k.store_link = -1;
    if (!cur.peek()) return true;  // no (more) blocks
    k.enabled = 1;

    if (cur.is_mov(SKR_S, SKR_X)) { cur.take(); k.save_s = 1; }
    if (cur.is_load(SKR_X)) {
        const skr_op* o = cur.take();
        if (o->b & 1) return false;
        k.sample_in = o->src;
    }
    if (cur.is_mov(SKR_B, SKR_P)) cur.take();
    else if (cur.is_load(SKR_B)) {
        const skr_op* o = cur.take();
        if (o->b & 1) return false;
        k.base_in = o->src;
    }

    const skr_op* o = cur.peek();
    if (!o) return false;
    bool pred_from_a = false;
    if (o->code == SKR_OP_ACC0) {
        k.kind = BK_ACC;
        bool head = true;
        while (cur.is(SKR_OP_ACC0) || cur.is(SKR_OP_ACC)) {
            const skr_op* t = cur.peek();
            if (t->code == SKR_OP_ACC0 && !head) return false;
            if (t->a == 0) {
                if (k.p_mode == 2) return false;  // tensor term after the trailing P term
                if (k.n_terms >= kMaxTerms) return false;
                k.term_in[k.n_terms] = t->src;
                k.terms[k.n_terms].c0 = (CT)t->c[0];
                ++k.n_terms;
            } else if (t->a == SKR_P + 1) {
                if (k.p_mode != 0) return false;
                k.p_mode = head ? 1 : 2;
                k.p_coef = (CT)t->c[0];
            } else {
                return false;
            }
            head = false;
            cur.take();
        }
        if (k.base_in >= 0) return false;  // ACC never reads B
        pred_from_a = true;
    } else if (o->code == SKR_OP_UNI || o->code == SKR_OP_UNIC || o->code == SKR_OP_ADDB) {
        k.kind = BK_UNI;
        int seen = 0;
        while (cur.is(SKR_OP_UNI)) {
            const skr_op* t = cur.take();
            if ((t->a != 0) != (seen == 0)) return false;
            if (k.n_terms >= kMaxTerms) return false;
            k.term_in[k.n_terms] = t->src;
            k.terms[k.n_terms].c0 = (CT)t->c[0];
            k.terms[k.n_terms].c1 = (CT)t->c[1];
            ++k.n_terms;
            ++seen;
        }
        if (cur.is(SKR_OP_UNIC)) {
            const skr_op* t = cur.take();
            if ((t->a != 0) != (seen == 0)) return false;
            k.p_mode = 1;
            k.p_coef = (CT)t->c[1];
            ++seen;
        }
        if (!cur.is(SKR_OP_ADDB)) return false;
        const skr_op* t = cur.take();
        k.empty_sum = t->a ? 1 : 0;
        if ((seen == 0) != (k.empty_sum != 0)) return false;
        pred_from_a = true;
    } else if (o->code == SKR_OP_DPM2) {
        const skr_op* t = cur.take();
        k.kind = BK_DPM2;
        k.n_terms = 1;
        k.term_in[0] = t->src;
        k.terms[0].c0 = (CT)t->c[0];
        k.terms[0].c1 = (CT)t->c[1];
        pred_from_a = true;
    } else if (o->code == SKR_OP_DPM3A) {
        const skr_op* a = cur.take();
        if (!cur.is(SKR_OP_DPM3B)) return false;
        const skr_op* b = cur.take();
        if (!cur.is(SKR_OP_DPM3C)) return false;
        const skr_op* c = cur.take();
        k.kind = BK_DPM3;
        k.n_terms = 2;
        k.term_in[0] = a->src;
        k.terms[0].c0 = (CT)a->c[0];
        k.term_in[1] = b->src;
        k.terms[1].c0 = (CT)b->c[0];
        k.terms[1].c1 = (CT)b->c[1];
        k.e0 = (CT)b->c[2];
        k.e1 = (CT)c->c[0];
        k.e2 = (CT)c->c[1];
        pred_from_a = true;
    }
    if (cur.is(SKR_OP_DIVA)) {
        if (k.kind != BK_ACC) return false;
        const skr_op* t = cur.take();
        k.has_div = 1;
        k.div = (CT)t->c[0];
    }
    if (!cur.is(SKR_OP_FWD)) return false;
    {
        const skr_op* t = cur.take();
        if (t->a == SKR_A) { if (!pred_from_a) return false; k.pred_is_p = 0; }
        else if (t->a == SKR_P) { if (pred_from_a) return false; k.pred_is_p = 1; }
        else return false;
        k.gamma = (CT)t->c[0];
        k.delta = (CT)t->c[1];
        if (t->b & 1) { k.has_noise = 1; k.noise_in = t->src; k.zeta = (CT)t->c[2]; }
        else if (t->b & 2) { k.has_noise = 2; k.noise_in = t->src; k.zeta = (CT)t->c[2]; }
    }
    if (cur.is_store(SKR_R)) k.store_r = cur.take()->dst;

    if (cur.is_mov(SKR_X, SKR_R)) { cur.take(); k.link = BL_X_FROM_R; }
    else if (cur.is(SKR_OP_BLEND)) {
        const skr_op* t = cur.take();
        if (t->a > 1 || !k.save_s) return false;
        k.link = t->a ? BL_BLEND_POW : BL_BLEND;
        k.l0 = (CT)t->c[0];
        k.l1 = (CT)t->c[1];
        k.lp = (CT)t->c[2];
        k.lq = (CT)t->c[3];
        if (cur.is_store(SKR_X)) k.store_link = cur.take()->dst;
    } else if (cur.is(SKR_OP_BACK)) {
        const skr_op* t = cur.take();
        if (t->b & 1) return false;
        k.link = BL_BACK;
        k.l0 = (CT)t->c[0];
        k.l1 = (CT)t->c[1];
        if (cur.is_store(SKR_P)) k.store_link = cur.take()->dst;
    }
    return true;
}

template <typename CT>
static bool block_reads_p(const BBlock<CT>& k) {
    return k.enabled && (k.pred_is_p || k.p_mode != 0 || (k.kind != BK_NONE && k.kind != BK_ACC && k.base_in < 0));
}

// Host-side reciprocals of every divisor in the descriptor (machine.cuh, div_uniform).
template <typename CT>
static void fill_reciprocals(BProgram<CT>& out) {
    bool fast = true;
    BHead<CT>& h = out.head;
    for (int c = 0; c < h.n_conv; ++c)
        if (h.conv_flags[c] & SKR_CONV_DIV) h.conv_r[c] = uniform_reciprocal(h.conv_c[c][2], &fast);
    for (int b = 0; b < 2; ++b) {
        BBlock<CT>& k = out.blk[b];
        if (!k.enabled) continue;
        if (k.has_div) k.div_r = uniform_reciprocal(k.div, &fast);
        if (k.link == BL_BACK) k.l1_r = uniform_reciprocal(k.l1, &fast);
        if (k.kind == BK_UNI)
            for (int t = 0; t < k.n_terms; ++t) k.terms[t].r0 = uniform_reciprocal(k.terms[t].c0, &fast);
    }
    out.fast_div = (fast && sizeof(CT) == 4) ? 1 : 0;
}

// Staged byte offsets of every operand (after in_off[] is known for the launch's tile size).
template <typename CT>
static void resolve_offsets(BProgram<CT>& k) {
    auto at = [&](int in) { return in >= 0 ? k.in_off[in] : 0u; };
    k.head.x_off = at(k.head.x_in);
    k.head.y_off = at(k.head.y_in);
    for (int b = 0; b < 2; ++b) {
        BBlock<CT>& blk = k.blk[b];
        blk.sample_off = at(blk.sample_in);
        blk.base_off = at(blk.base_in);
        blk.noise_off = blk.has_noise == 1 ? at(blk.noise_in) : 0u;
        for (int t = 0; t < blk.n_terms; ++t) blk.terms[t].off = at(blk.term_in[t]);
    }
}

// Returns true when `p` matches the skeleton; fills head/blk of `out`.
template <typename CT>
static bool parse_block_program(const skr_program* p, BProgram<CT>& out) {
    OpCursor cur{p, 0};
    BHead<CT>& h = out.head;
    memset(&h, 0, sizeof(h));
    h.x_in = h.y_in = h.store_p = h.store_p2 = -1;

    if (cur.is_load(SKR_X)) {
        const skr_op* o = cur.take();
        if (o->b & 1) return false;
        h.x_in = o->src;
    }
    if (cur.is_load(SKR_P)) {
        const skr_op* o = cur.take();
        h.y_in = o->src;
        h.neg = o->b & 1;
    } else if (cur.is(SKR_OP_CONV) && cur.peek()->b == 0) {
        const skr_op* o = cur.take();
        h.y_in = o->src;
        h.conv_flags[0] = o->a;
        for (int j = 0; j < 3; ++j) h.conv_c[0][j] = (CT)o->c[j];
        h.n_conv = 1;
    }
    while (cur.is(SKR_OP_CONV) && cur.peek()->b == 1) {
        if (h.n_conv >= 2 || h.y_in < 0) return false;
        const skr_op* o = cur.take();
        h.conv_flags[h.n_conv] = o->a;
        for (int j = 0; j < 3; ++j) h.conv_c[h.n_conv][j] = (CT)o->c[j];
        ++h.n_conv;
    }
    for (int c = 0; c < h.n_conv; ++c)
        if ((h.conv_flags[c] & SKR_CONV_USE_X) && h.x_in < 0) return false;
    if (cur.is_store(SKR_P)) h.store_p = cur.take()->dst;
    if (cur.is_store(SKR_P)) h.store_p2 = cur.take()->dst;

    if (!parse_block<CT>(cur, out.blk[0])) return false;
    if (!parse_block<CT>(cur, out.blk[1])) return false;
    if (cur.peek()) return false;  // trailing ops the skeleton cannot express
    fill_reciprocals<CT>(out);
    if (out.blk[0].enabled && h.x_in < 0 && out.blk[0].sample_in < 0) return false;  // X must come from somewhere
    if ((block_reads_p(out.blk[0]) || block_reads_p(out.blk[1]) || h.store_p >= 0) && h.y_in < 0) return false;
    return out.blk[0].enabled || h.store_p >= 0;
}

}  // namespace skr
