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
// eight state vectors live in fixed registers.  A dedicated producer warp issues the bulk copies
// (one lane per input tensor) and runs ahead of the eight consumer warps through a full/empty
// mbarrier ring; consumers never synchronise with each other.
//
// Arithmetic is identical to the interpreter (machine.cuh): individually rounded ops in the
// reference's order.  A program that does not fit the skeleton is executed by the interpreter.
#pragma once

#include "machine.cuh"

namespace skr {

constexpr int kMaxTerms = 36;
constexpr int kProducerThreads = 32;

enum BlockKind : uint8_t { BK_NONE = 0, BK_ACC = 1, BK_UNI = 2, BK_DPM2 = 3, BK_DPM3 = 4 };
enum BlockLink : uint8_t { BL_NONE = 0, BL_X_FROM_R = 1, BL_BLEND = 2, BL_BACK = 3 };

template <typename CT>
struct BTerm {
    CT c0, c1;
    int32_t in;  // input index
};

template <typename CT>
struct BBlock {
    uint8_t enabled, kind, save_s, p_mode;       // p_mode: ACC 1 = P first, 2 = P last; UNI 1 = UniC term
    uint8_t has_div, pred_is_p, has_noise, link;
    uint8_t n_terms, empty_sum, pad0, pad1;
    int8_t sample_in, base_in, noise_in, store_r, store_link, pad2, pad3, pad4;  // -1 = not used
    CT p_coef, div, gamma, delta, zeta, l0, l1, e0, e1, e2;
    BTerm<CT> terms[kMaxTerms];
};

template <typename CT>
struct BHead {
    int8_t x_in, y_in, store_p, n_conv;
    uint8_t neg, conv_flags[2], pad;
    CT conv_c[2][3];
};

template <typename CT>
struct BProgram {
    int64_t numel;
    int32_t n_inputs, stages;
    uint32_t stage_bytes, use_tma;
    const void* in_ptr[SKR_MAX_INPUTS];
    void* out_ptr[SKR_MAX_OUTPUTS];
    uint32_t in_off[SKR_MAX_INPUTS];
    uint8_t in_dtype[SKR_MAX_INPUTS];
    uint8_t out_dtype[SKR_MAX_OUTPUTS];
    BHead<CT> head;
    BBlock<CT> blk[2];
};

template <typename CT, bool DIRECT>
struct Fetcher {
    const BProgram<CT>& prog;
    const unsigned char* stage;
    int tid;
    int64_t first;
    __device__ __forceinline__ void operator()(int in, CT (&v)[kVec]) const {
        if constexpr (DIRECT) fetch_direct<CT>(prog.in_ptr[in], prog.in_dtype[in], first, prog.numel, v);
        else fetch_staged<CT>(stage, prog.in_off[in], prog.in_dtype[in], tid, v);
    }
};

template <typename CT, bool DIRECT>
__device__ __forceinline__ void run_block_tile(const BProgram<CT>& prog, int64_t tile, const unsigned char* stage, int tid) {
    using Ar = Arith<CT>;
    const int64_t first = tile * kTile + (int64_t)tid * kVec;
    const Fetcher<CT, DIRECT> fetch{prog, stage, tid, first};
    const int64_t numel = prog.numel;

    CT X[kVec], P[kVec], B[kVec], A[kVec], S[kVec], R[kVec];
#pragma unroll
    for (int j = 0; j < kVec; ++j) X[j] = P[j] = B[j] = A[j] = S[j] = R[j] = (CT)0;

    // ---- head ---------------------------------------------------------------------------------
    const BHead<CT>& h = prog.head;
    if (h.x_in >= 0) fetch(h.x_in, X);
    if (h.y_in >= 0) {
        fetch(h.y_in, P);
        if (h.neg) {
#pragma unroll
            for (int j = 0; j < kVec; ++j) P[j] = -P[j];
        }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c < h.n_conv) {
            const int f = h.conv_flags[c];
            const CT c0 = h.conv_c[c][0], c1 = h.conv_c[c][1], c2 = h.conv_c[c][2];
#pragma unroll
            for (int j = 0; j < kVec; ++j) {
                CT v;
                if (f & SKR_CONV_USE_X) {
                    const CT lhs = (f & SKR_CONV_MUL_X) ? Ar::mul(c0, X[j]) : X[j];
                    const CT rhs = (f & SKR_CONV_MUL_Y) ? Ar::mul(c1, P[j]) : P[j];
                    v = Ar::sub(lhs, rhs);
                } else {
                    v = (f & SKR_CONV_MUL_Y) ? Ar::mul(P[j], c1) : P[j];
                }
                P[j] = (f & SKR_CONV_DIV) ? Ar::div(v, c2) : v;
            }
        }
    }
    if (h.store_p >= 0) store_vec<CT, DIRECT>(prog.out_ptr[h.store_p], prog.out_dtype[h.store_p], first, numel, P);

    // ---- blocks -------------------------------------------------------------------------------
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const BBlock<CT>& k = prog.blk[b];
        if (!k.enabled) continue;
        if (k.save_s) {
#pragma unroll
            for (int j = 0; j < kVec; ++j) S[j] = X[j];
        }
        if (k.sample_in >= 0) fetch(k.sample_in, X);
        if (k.base_in >= 0) fetch(k.base_in, B);
        else {
#pragma unroll
            for (int j = 0; j < kVec; ++j) B[j] = P[j];
        }

        const int n_terms = k.n_terms;
        CT in[kVec];
        switch (k.kind) {
            case BK_ACC: {
                int t = 0;
                if (k.p_mode == 1 || n_terms == 0) {
#pragma unroll
                    for (int j = 0; j < kVec; ++j) A[j] = Ar::add((CT)0, Ar::mul(P[j], k.p_coef));
                } else {
                    fetch(k.terms[0].in, in);
                    const CT c = k.terms[0].c0;
#pragma unroll
                    for (int j = 0; j < kVec; ++j) A[j] = Ar::add((CT)0, Ar::mul(in[j], c));
                    t = 1;
                }
                for (; t < n_terms; ++t) {
                    fetch(k.terms[t].in, in);
                    const CT c = k.terms[t].c0;
#pragma unroll
                    for (int j = 0; j < kVec; ++j) A[j] = Ar::add(A[j], Ar::mul(in[j], c));
                }
                if (k.p_mode == 2 && n_terms > 0) {
#pragma unroll
                    for (int j = 0; j < kVec; ++j) A[j] = Ar::add(A[j], Ar::mul(P[j], k.p_coef));
                }
            } break;
            case BK_UNI: {
                for (int t = 0; t < n_terms; ++t) {
                    fetch(k.terms[t].in, in);
                    const CT rk = k.terms[t].c0, rho = k.terms[t].c1;
#pragma unroll
                    for (int j = 0; j < kVec; ++j) {
                        const CT term = Ar::mul(Ar::div(Ar::sub(in[j], B[j]), rk), rho);
                        A[j] = Ar::add(t == 0 ? (CT)0 : A[j], term);
                    }
                }
                if (k.p_mode == 1) {
#pragma unroll
                    for (int j = 0; j < kVec; ++j) {
                        const CT term = Ar::mul(Ar::sub(P[j], B[j]), k.p_coef);
                        A[j] = Ar::add(n_terms == 0 ? (CT)0 : A[j], term);
                    }
                }
#pragma unroll
                for (int j = 0; j < kVec; ++j) A[j] = Ar::add(B[j], k.empty_sum ? (CT)0 : A[j]);
            } break;
            case BK_DPM2: {
                fetch(k.terms[0].in, in);
                const CT inv_r = k.terms[0].c0, half = k.terms[0].c1;
#pragma unroll
                for (int j = 0; j < kVec; ++j) A[j] = Ar::add(B[j], Ar::mul(half, Ar::mul(inv_r, Ar::sub(B[j], in[j]))));
            } break;
            case BK_DPM3: {
                CT in2[kVec];
                fetch(k.terms[0].in, in);
                fetch(k.terms[1].in, in2);
                const CT inv_r = k.terms[0].c0, inv_r2 = k.terms[1].c0, mix = k.terms[1].c1;
                const CT inv_sum = k.e0, w1 = k.e1, w2 = k.e2;
#pragma unroll
                for (int j = 0; j < kVec; ++j) {
                    const CT d10 = Ar::mul(inv_r, Ar::sub(B[j], in[j]));
                    const CT d11 = Ar::mul(inv_r2, Ar::sub(in[j], in2[j]));
                    const CT d = Ar::sub(d10, d11);
                    const CT d1 = Ar::add(d10, Ar::mul(mix, d));
                    const CT d2 = Ar::mul(inv_sum, d);
                    A[j] = Ar::add(Ar::add(B[j], Ar::mul(w1, d1)), Ar::mul(w2, d2));
                }
            } break;
            default: break;
        }
        if (k.has_div) {
#pragma unroll
            for (int j = 0; j < kVec; ++j) A[j] = Ar::div(A[j], k.div);
        }
        if (k.has_noise) fetch(k.noise_in, in);
#pragma unroll
        for (int j = 0; j < kVec; ++j) {
            const CT pred = k.pred_is_p ? P[j] : A[j];
            CT v = Ar::add((CT)0, Ar::mul(X[j], k.gamma));
            v = Ar::add(v, Ar::mul(pred, k.delta));
            if (k.has_noise) v = Ar::add(v, Ar::mul(in[j], k.zeta));
            R[j] = v;
        }
        if (k.store_r >= 0) store_vec<CT, DIRECT>(prog.out_ptr[k.store_r], prog.out_dtype[k.store_r], first, numel, R);

        if (k.link == BL_X_FROM_R) {
#pragma unroll
            for (int j = 0; j < kVec; ++j) X[j] = R[j];
        } else if (k.link == BL_BLEND) {
#pragma unroll
            for (int j = 0; j < kVec; ++j) X[j] = Ar::add(Ar::mul(S[j], k.l0), Ar::mul(R[j], k.l1));
            if (k.store_link >= 0) store_vec<CT, DIRECT>(prog.out_ptr[k.store_link], prog.out_dtype[k.store_link], first, numel, X);
        } else if (k.link == BL_BACK) {
#pragma unroll
            for (int j = 0; j < kVec; ++j) P[j] = Ar::div(Ar::sub(R[j], Ar::mul(X[j], k.l0)), k.l1);
            if (k.store_link >= 0) store_vec<CT, DIRECT>(prog.out_ptr[k.store_link], prog.out_dtype[k.store_link], first, numel, P);
        }
    }
}

template <typename CT>
__global__ void __launch_bounds__(kThreads + kProducerThreads) block_kernel(const __grid_constant__ BProgram<CT> prog) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t numel = prog.numel;
    const int64_t n_full = prog.use_tma ? numel / kTile : 0;
    const int64_t n_tiles = (numel + kTile - 1) / kTile;
    const int stages = prog.stages;
    const uint32_t stage_bytes = prog.stage_bytes;
    const int64_t mine = n_full > (int64_t)blockIdx.x ? (n_full - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const bool producer = warp == kThreads / 32;

    if (mine > 0) {
        if (tid == 0) {
            for (int s = 0; s < stages; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], kThreads / 32);
            }
            fence_barrier_init();
        }
        __syncthreads();

        if (producer) {
            // one lane per input tensor; lanes beyond n_inputs idle (SKR_MAX_INPUTS == 32 == warp size)
            const bool active = lane < prog.n_inputs;
            const uint32_t esize = active ? dtype_size(prog.in_dtype[lane]) : 0u;
            const unsigned char* src = active ? reinterpret_cast<const unsigned char*>(prog.in_ptr[lane]) : nullptr;
            const uint32_t off = active ? prog.in_off[lane] : 0u;
            for (int64_t k = 0; k < mine; ++k) {
                const int s = (int)(k % stages);
                if (k >= stages) mbar_wait(&empty_bar[s], (uint32_t)(((k / stages) - 1) & 1));
                if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
                __syncwarp();
                if (active) {
                    const int64_t tile = blockIdx.x + k * (int64_t)gridDim.x;
                    tma_load_1d(smem + (size_t)s * stage_bytes + off, src + (size_t)tile * kTile * esize, kTile * esize, &full_bar[s]);
                }
            }
        } else {
            for (int64_t k = 0; k < mine; ++k) {
                const int s = (int)(k % stages);
                mbar_wait(&full_bar[s], (uint32_t)((k / stages) & 1));
                run_block_tile<CT, false>(prog, blockIdx.x + k * (int64_t)gridDim.x, smem + (size_t)s * stage_bytes, tid);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
            }
        }
    }
    if (!producer) {
        for (int64_t tile = n_full + blockIdx.x; tile < n_tiles; tile += gridDim.x) run_block_tile<CT, true>(prog, tile, nullptr, tid);
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
static bool parse_block(OpCursor& cur, BBlock<CT>& k, bool first_block) {
    memset(&k, 0, sizeof(k));
    k.sample_in = k.base_in = k.noise_in = k.store_r = k.store_link = -1;
    if (!cur.peek()) return true;  // no (more) blocks
    k.enabled = 1;

    if (cur.is_mov(SKR_S, SKR_X)) { cur.take(); k.save_s = 1; }
    if (cur.is_load(SKR_X)) {
        const skr_op* o = cur.take();
        if (o->b & 1) return false;
        k.sample_in = (int8_t)o->src;
    }
    bool base_set = false;
    if (cur.is_mov(SKR_B, SKR_P)) { cur.take(); base_set = true; }
    else if (cur.is_load(SKR_B)) {
        const skr_op* o = cur.take();
        if (o->b & 1) return false;
        k.base_in = (int8_t)o->src;
        base_set = true;
    }
    (void)base_set;
    (void)first_block;

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
                k.terms[k.n_terms].in = t->src;
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
        pred_from_a = true;
    } else if (o->code == SKR_OP_UNI || o->code == SKR_OP_UNIC || o->code == SKR_OP_ADDB) {
        k.kind = BK_UNI;
        int seen = 0;
        while (cur.is(SKR_OP_UNI)) {
            const skr_op* t = cur.take();
            if ((t->a != 0) != (seen == 0)) return false;
            if (k.n_terms >= kMaxTerms) return false;
            k.terms[k.n_terms].in = t->src;
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
        k.terms[0].in = t->src;
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
        k.terms[0].in = a->src;
        k.terms[0].c0 = (CT)a->c[0];
        k.terms[1].in = b->src;
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
        if (t->b & 1) { k.has_noise = 1; k.noise_in = (int8_t)t->src; k.zeta = (CT)t->c[2]; }
    }
    if (cur.is_store(SKR_R)) k.store_r = (int8_t)cur.take()->dst;

    if (cur.is_mov(SKR_X, SKR_R)) { cur.take(); k.link = BL_X_FROM_R; }
    else if (cur.is(SKR_OP_BLEND)) {
        const skr_op* t = cur.take();
        if (t->a != 0 || !k.save_s) return false;
        k.link = BL_BLEND;
        k.l0 = (CT)t->c[0];
        k.l1 = (CT)t->c[1];
        if (cur.is_store(SKR_X)) k.store_link = (int8_t)cur.take()->dst;
    } else if (cur.is(SKR_OP_BACK)) {
        const skr_op* t = cur.take();
        if (t->b & 1) return false;
        k.link = BL_BACK;
        k.l0 = (CT)t->c[0];
        k.l1 = (CT)t->c[1];
        if (cur.is_store(SKR_P)) k.store_link = (int8_t)cur.take()->dst;
    }
    return true;
}

// Returns true when `p` matches the skeleton; fills head/blk of `out`.
template <typename CT>
static bool parse_block_program(const skr_program* p, BProgram<CT>& out) {
    OpCursor cur{p, 0};
    BHead<CT>& h = out.head;
    memset(&h, 0, sizeof(h));
    h.x_in = h.y_in = h.store_p = -1;

    if (cur.is_load(SKR_X)) {
        const skr_op* o = cur.take();
        if (o->b & 1) return false;
        h.x_in = (int8_t)o->src;
    }
    if (cur.is_load(SKR_P)) {
        const skr_op* o = cur.take();
        h.y_in = (int8_t)o->src;
        h.neg = o->b & 1;
    } else if (cur.is(SKR_OP_CONV) && cur.peek()->b == 0) {
        const skr_op* o = cur.take();
        h.y_in = (int8_t)o->src;
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
    if (cur.is_store(SKR_P)) h.store_p = (int8_t)cur.take()->dst;

    if (!parse_block<CT>(cur, out.blk[0], true)) return false;
    if (!parse_block<CT>(cur, out.blk[1], false)) return false;
    if (cur.peek()) return false;  // trailing ops the skeleton cannot express
    // a block that uses X needs it to come from somewhere
    if (out.blk[0].enabled && h.x_in < 0 && out.blk[0].sample_in < 0) return false;
    if ((out.blk[0].enabled && (out.blk[0].kind == BK_NONE || out.blk[0].base_in < 0)) && h.y_in < 0) return false;
    return out.blk[0].enabled || h.store_p >= 0;
}

}  // namespace skr
