/*
 * skrample_b200 - C ABI of the B200-native sampler-step library.
 *
 * The reference (Beinsezii/skrample) is pure Python and has no FFI: its hot
 * path is a chain of ATen elementwise ops issued from
 *   skrample/sampling/models.py:53-83      (forward / backward)
 *   skrample/sampling/models.py:92-212     (to_x / from_x conversions)
 *   skrample/sampling/structured.py:167-577 (Euler, DPM, Adams, UniP, UniPC, SPC)
 *   skrample/sampling/functional.py:55-105  (step_tableau)
 *   skrample/common.py:32-40               (Point.add_noise / remove_noise)
 *   skrample/pytorch/noise.py:36-425       (Random / Offset / Pyramid / Colored)
 *   skrample/sampling/functional.py:197-214 (adaptive error norms)
 * Every entry point below replaces one of those call sites with ONE launch of a
 * hand-written sm_100a kernel (the composite noise generators: a few).  Plain
 * pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns int: 0 = ok, < 0 = bad argument (see SKR_E_*),
 *     > 0 = a cudaError_t from the launch.  skr_last_error() gives the text
 *     (thread-local).  No C++ exception crosses this boundary.
 *   - launches are asynchronous on the caller's stream (`stream` is a
 *     cudaStream_t passed as void*), on the current device, and never
 *     synchronise, allocate, free or retain device memory.
 *   - tensors are dense, contiguous, element-count `numel`; all tensors of one
 *     call have the same numel.  16-byte aligned bases take the TMA fast path,
 *     anything else a slower but correct element-wise path.
 *   - scalars are float64; the library rounds them to the compute type
 *     (fp32, or fp64 when any tensor is fp64) exactly once, like torch does
 *     with a Python scalar operand.
 */
#ifndef SKRAMPLE_B200_H
#define SKRAMPLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKR_VERSION 1

#define SKR_MAX_OPS 64
#define SKR_MAX_INPUTS 32
#define SKR_MAX_OUTPUTS 8
#define SKR_MAX_DIMS 8
#define SKR_MAX_PHILOX 2
#define SKR_MAX_PHILOX_ITEMS 256
#define SKR_MAX_LEVELS 16
#define SKR_BROWNIAN_MAX_DEPTH 40
/* Accumulators of grid-wide sums (`moments`, `sums` below): device double[SKR_MOMENTS_DOUBLES], zeroed before first use.
 * [0], [1] hold the two sums; the rest is scratch for a summation in a fixed order (bit-reproducible results). */
#define SKR_MOMENT_BLOCKS 2048
#define SKR_MOMENTS_DOUBLES (4 + 2 * SKR_MOMENT_BLOCKS)

/* element types */
enum { SKR_F32 = 0, SKR_F64 = 1, SKR_BF16 = 2, SKR_F16 = 3 };

/* error codes */
enum {
    SKR_E_NULL = -1,      /* null program / pointer            */
    SKR_E_RANGE = -2,     /* count or index out of range        */
    SKR_E_DTYPE = -3,     /* unknown element type               */
    SKR_E_OPCODE = -4,    /* unknown op code                    */
    SKR_E_SHAPE = -5,     /* inconsistent shape description     */
    SKR_E_UNSUPPORTED = -6
};

/* registers of the step machine (each holds one value per latent element) */
enum { SKR_X = 0, SKR_P = 1, SKR_B = 2, SKR_A = 3, SKR_S = 4, SKR_R = 5, SKR_T = 6, SKR_U = 7 };

/*
 * Op codes.  "in" is the op's input tensor element, c0..c3 its scalars.  Each
 * binary operation is individually rounded (no FMA contraction) in the order
 * written, which is the reference's evaluation order.
 */
enum {
    SKR_OP_END = 0,
    SKR_OP_LOAD = 1,   /* reg[a] = (b&1) ? -in : in                                             */
    SKR_OP_MOV = 2,    /* reg[a] = reg[b]                                                       */
    SKR_OP_STORE = 3,  /* out[dst] = round_to_dtype(reg[a])                                     */
    SKR_OP_CONV = 4,   /* P = conv(X, y), y = in (b==0) or P (b==1); a = SKR_CONV_* flags:
                          USE_X: v = (MUL_X ? c0*X : X) - (MUL_Y ? c1*y : y) else v = MUL_Y ? y*c1 : y;
                          DIV: v = v / c2                           models.py:92-212            */
    SKR_OP_ACC0 = 5,   /* A = 0 + s*c0, s = in (a==0) or reg[a-1]    math.sumprod head          */
    SKR_OP_ACC = 6,    /* A = A + s*c0                                                          */
    SKR_OP_DIVA = 7,   /* A = A / c0                                 functional.py:84           */
    SKR_OP_UNI = 8,    /* A = (a ? 0 : A) + ((in - B)/c0)*c1         structured.py:390-428      */
    SKR_OP_UNIC = 9,   /* A = (a ? 0 : A) + (P - B)*c1               structured.py:403          */
    SKR_OP_ADDB = 10,  /* A = B + (a ? 0 : A)                        structured.py:428          */
    SKR_OP_DPM2 = 11,  /* A = B + c1*(c0*(B - in))                   structured.py:243,269      */
    SKR_OP_DPM3A = 12, /* T = in; U = c0*(B - in)                    structured.py:243          */
    SKR_OP_DPM3B = 13, /* d11 = c0*(T - in); d = U - d11; T = U + c1*d; U = c2*d   :254-256     */
    SKR_OP_DPM3C = 14, /* A = (B + c0*T) + c1*U                      structured.py:268          */
    SKR_OP_FWD = 15,   /* R = ((0 + X*c0) + reg[a]*c1) [+ n*c2]; n = in (b&1) or a normal drawn in the
                          kernel from philox[src] (b&2)                     models.py:53-67      */
    SKR_OP_BACK = 16,  /* P = ((R - X*c0) [- in*c2 if b&1]) / c1           models.py:69-83      */
    SKR_OP_BLEND = 17, /* a==0: X = S*c0 + R*c1; a==1: signed-power blend, c2 = pw, c3 = 1/pw
                                                                      structured.py:568-575     */
    SKR_OP_AXPBY = 18, /* a==0: R = X*c0 + in*c1; a==1: R = (X - in*c0)/c1   common.py:32-40    */
    SKR_OP__COUNT = 19
};

enum { SKR_CONV_USE_X = 1, SKR_CONV_MUL_X = 2, SKR_CONV_MUL_Y = 4, SKR_CONV_DIV = 8 };

typedef struct skr_op {
    uint8_t code;
    uint8_t a;
    uint8_t b;
    uint8_t reserved;
    int16_t src; /* input index, -1 if the op reads no tensor  */
    int16_t dst; /* output index, -1 if the op writes no tensor */
    double c[4];
} skr_op;

typedef struct skr_tensor {
    void* ptr;     /* device pointer */
    int32_t dtype; /* SKR_F32 ...    */
    int32_t reserved;
} skr_tensor;

/*
 * A noise tensor that is never materialised: element e of batch item i is the standard normal of Philox
 * stream (seed[i], stream[i]) at counter e / 4 - bit-identical to what skr_noise_fill writes for that item
 * (noise.py:36-42 + BatchTensorNoise noise.py:438-466: one generator per batch item).
 */
typedef struct skr_philox {
    uint64_t seed[SKR_MAX_PHILOX_ITEMS];
    uint64_t stream[SKR_MAX_PHILOX_ITEMS];
    int64_t item_numel; /* elements per batch item; numel == n_items * item_numel */
    int32_t n_items;    /* 1..SKR_MAX_PHILOX_ITEMS */
    int32_t dtype;      /* storage type of the tensor this stands for: SKR_BF16 / SKR_F16 round every normal to that
                           type before it is used, exactly what reading the filled tensor would give; SKR_F32 (0) and
                           SKR_F64 use the fp32 normal as drawn */
    /* Offset noise (noise.py:84-113, kept axes leading): when offset_inner > 0 every run of offset_inner consecutive
       elements of an item shares one more normal - stream[i] + 1 of the same seed at counter = run index - added to
       each of its elements times offset_scale (strength^2): what skr_noise_fill writes with that skr_offset.
       item_numel must be a multiple of offset_inner.  0: plain Random. */
    int64_t offset_inner;
    float offset_scale;
    int32_t reserved;
} skr_philox;

typedef struct skr_program {
    int32_t n_ops;
    int32_t n_inputs;
    int32_t n_outputs;
    int32_t n_philox;
    skr_op ops[SKR_MAX_OPS];
    skr_tensor inputs[SKR_MAX_INPUTS];
    skr_tensor outputs[SKR_MAX_OUTPUTS];
    skr_philox philox[SKR_MAX_PHILOX];
} skr_program;

/* Library identification / diagnostics. */
int skr_version(void);
const char* skr_last_error(void);
/* Number of kernels this library has launched in this process (bench bookkeeping). */
int64_t skr_launch_count(void);
/* Same, split by kernel: kind 0 = structured block kernel (fast path), 1 = interpreter (general path),
 * 2 = noise kernels. */
int64_t skr_launch_count_kind(int32_t kind);

/*
 * The fused solver step: run `program` over `numel` elements as ONE kernel.
 * Replaces the whole per-step ATen op chain of a structured sampler
 * (structured.py:167-577), an RK stage (functional.py:80-105), a model
 * conversion (models.py:215-239) or forward/backward (models.py:53-83).
 */
int skr_program_launch(const skr_program* program, int64_t numel, void* stream);

/*
 * Which kernel would run `program`: 0 = structured block kernel, 1 = interpreter, < 0 = error.
 * Pure host logic (no device needed); used by tests and tooling.
 */
int skr_program_classify(const skr_program* program);

/*
 * Like skr_program_classify, and writes a one-line description of the parsed step into `text`
 * (NUL-terminated, at most `capacity` bytes): compute type, the compiled kernel shape the step
 * would run with ("any" = generic instantiation) and the control fields of head and blocks.
 * Uses only the dtypes of the tensors, never the pointers.  Pure host logic; tests and tooling.
 */
int skr_program_describe(const skr_program* program, char* text, int32_t capacity);

/*
 * Step plans: the same launch with the host work done once.
 *
 * The reference re-derives every step from scratch (schedule points, dataclass churn, one ATen call per op:
 * structured.py:33-34,137-149); a sampling loop takes the SAME step - same ops, same scalars, same tensor types -
 * once per trajectory, so everything but the tensor addresses can be prepared ahead.  skr_plan_create validates
 * `program`, recognises the step's structure and chooses the kernel instantiation (only the ops, the counts and the
 * dtypes of program->inputs / outputs are read; pointers are ignored).  skr_plan_launch binds addresses and launches:
 * `tensors` holds the n_inputs input pointers followed by the n_outputs output pointers, in the program's order and
 * of the dtypes given at creation; `draws` = program->n_philox key tables (NULL when the step draws no noise).
 * A plan is immutable: any number of threads may launch it concurrently.  Same results, bit for bit, as
 * skr_program_launch on the same program.
 */
typedef struct skr_plan skr_plan;
int skr_plan_create(const skr_program* program, skr_plan** plan);
int skr_plan_launch(const skr_plan* plan, const void* const* tensors, int64_t numel, const skr_philox* draws, void* stream);
void skr_plan_destroy(skr_plan* plan);
/* 0 = structured block kernel, 1 = interpreter (what skr_program_classify says of the program), < 0 = error. */
int skr_plan_kind(const skr_plan* plan);
/* Name of the compiled kernel shape the plan launches ("any" = generic instantiation, "interpreter"). */
const char* skr_plan_shape(const skr_plan* plan);

/*
 * Development switches (SKR_FORCE_INTERP, SKR_NO_PINNED, SKR_IN_MODE, SKR_STAGES, SKR_CTAS) are read from the
 * environment once per process; tests and A/B tooling that change them afterwards call this to re-read them.
 * Existing plans keep the kernel they were created with.
 */
void skr_reload_env(void);

/*
 * Arithmetic of the fused step, process-wide; plans keep the mode they were created under.
 *   0  exact (default): every product, sum and quotient individually rounded in the reference's order - fp32 / fp64
 *      results are bit-identical to the reference's torch-CPU path;
 *   1  contracted: the divided-difference and weighted-sum steps (UniP / UniPC / SPC / Adams on fp32 compute) fuse
 *      a*b + c into one multiply-add and divide by multiplying with the reciprocal: a few fp32 ulp per step (inside the
 *      1e-5 relative per step the north star asks for), ~30% fewer instructions where the step is issue-bound.
 * The environment variable SKR_ARITH=contracted selects mode 1 at start-up.
 */
int skr_set_arithmetic(int32_t mode);
int skr_get_arithmetic(void);

/*
 * Point.add_noise / remove_noise (common.py:32-40):
 *   remove == 0: out = sample*alpha + noise*sigma
 *   remove != 0: out = (sample - noise*sigma) / alpha
 */
int skr_axpby(const void* sample, const void* noise, void* out, int32_t dtype, int64_t numel, double sigma,
              double alpha, int32_t remove, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Noise generation (skrample/pytorch/noise.py).  Counter-based Philox4x32-10: the value at element e of
 * stream (seed, stream) depends on nothing else, so shards of a batch reproduce the unsharded values.
 */

/* Offset term of noise.py:104-113: out += normal(offset stream)[index reduced to the kept axes] * scale */
typedef struct skr_offset {
    int32_t ndim;
    int32_t reserved;
    int64_t shape[SKR_MAX_DIMS]; /* full shape of the unit tensor (product == numel)          */
    int32_t keep[SKR_MAX_DIMS];  /* 1: the offset varies along this axis, 0: broadcast         */
    uint64_t stream;             /* Philox stream of the offset draw                           */
    double scale;                /* strength^2                                                 */
} skr_offset;

/*
 * Random / Offset (noise.py:73-74,104-113): out[e] = normal(seed, stream)[e] (+ offset term).
 * `moments` (optional, device double[SKR_MOMENTS_DOUBLES], zeroed) accumulates sum and sum of squares of the values written.
 */
int skr_noise_fill(void* out, int32_t dtype, int64_t numel, uint64_t seed, uint64_t stream, const skr_offset* offset,
                   double* moments, void* cuda_stream);

/*
 * The same for a whole batch in one launch: item i (keys->item_numel elements) is the normal stream
 * (keys->seed[i], keys->stream[i]) - BatchTensorNoise.generate without the per-item launches and the stack copy
 * (noise.py:445-446).  numel = keys->n_items * keys->item_numel.
 */
int skr_noise_fill_batch(void* out, int32_t dtype, const skr_philox* keys, void* cuda_stream);

/*
 * Brownian (noise.py:210-252; the reference delegates to torchsde.BrownianInterval(t0=0, t1=1, entropy=seed,
 * tol=1/(10*max_steps))): out = (W(t1) - W(t0)) * out_scale for the Brownian path W keyed by `seed` alone, so the
 * same (seed, t0, t1) gives the same tensor on every call, increments over adjoining intervals add up and
 * increments over disjoint intervals are independent.  W is a Levy bridge tree over dyadic intervals of 0..1,
 * `depth` levels deep (leaf width 2^-depth plays torchsde's `tol`), node draws = Philox streams 1<<63 | heap
 * index, one exact bridge draw inside the leaf.  0 <= t0 < t1 <= 1, 1 <= depth <= SKR_BROWNIAN_MAX_DEPTH.
 * Values are this library's own (torchsde is not available to pin against): parity is statistical.
 */
int skr_noise_brownian(void* out, int32_t dtype, int64_t numel, uint64_t seed, double t0, double t1, int32_t depth,
                       double out_scale, void* cuda_stream);

/*
 * The same for a batch in one launch: item i (item_numel elements, starting at element i * item_numel of `out`)
 * is the path of seeds[i] - BatchTensorNoise over per-item Brownian generators (noise.py:445-446) without the
 * per-item launches.  1 <= n_items <= SKR_MAX_PHILOX_ITEMS; `seeds` is a host array.
 */
int skr_noise_brownian_batch(void* out, int32_t dtype, const uint64_t* seeds, int32_t n_items, int64_t item_numel,
                             double t0, double t1, int32_t depth, double out_scale, void* cuda_stream);

/* sum / sum^2 of a tensor into device double[SKR_MOMENTS_DOUBLES] (zeroed), for Tensor.std() (noise.py:207,365,401). */
int skr_noise_moments(const void* in, int32_t dtype, int64_t numel, double* moments, void* cuda_stream);

/*
 * Error norms of an embedded Runge-Kutta pair in one pass (FunctionalAdaptive.mae/.mse applied to (low, high) and
 * (0, high), functional.py:197-214, used by RKMoire.sample_model functional.py:437-441):
 *   sums[0] += sum |low - high|^power,  sums[1] += sum |high|^power   (power = 1 or 2; device double[SKR_MOMENTS_DOUBLES], zeroed)
 */
int skr_error_norms(const void* low, const void* high, int32_t dtype, int64_t numel, int32_t power, double* sums,
                    void* cuda_stream);

/*
 * out = in * s, s = numerator [* std(num_moments, num_count)] [/ std(moments, count)], each factor applied when
 * its pointer is given (unbiased std from the first two doubles of an accumulator); s stays 1 when the denominator std
 * is <= min_std.  In place allowed; casts in_dtype -> out_dtype.  (noise.py:207, 369, 402-405)
 */
int skr_noise_scale(const void* in, int32_t in_dtype, void* out, int32_t out_dtype, int64_t numel, double numerator,
                    const double* num_moments, int64_t num_count, const double* moments, int64_t count, double min_std,
                    void* cuda_stream);

typedef struct skr_pyramid_level {
    uint64_t stream;     /* Philox stream of this level's normal draw                               */
    const float* buffer; /* non-null: read the level (fp32, level shape, row-major) instead          */
    int64_t extent[2];   /* extents of the resized axes at this level, in axis order                  */
    double weight;       /* strength^level; 0 skips the level                                         */
} skr_pyramid_level;

typedef struct skr_pyramid {
    int32_t ndim;
    int32_t n_levels;
    int64_t shape[SKR_MAX_DIMS];  /* unit tensor shape                                                */
    int32_t masked[SKR_MAX_DIMS]; /* 1: axis is resized by the pyramid (1 or 2 axes)                  */
    uint64_t seed;
    uint64_t base_stream;         /* the plain randn(shape) term                                      */
    const float* base_buffer;     /* non-null: supplied base draw                                     */
    float* scratch;               /* optional, numel floats (may alias base_buffer): see below        */
    float* levels_scratch;        /* optional work area of levels_scratch_floats floats.  Weighted levels smaller
                                     than the unit that come without a buffer are drawn into it by one launch
                                     (each rounded up to a multiple of 4 floats) instead of inside the composition;
                                     when the resized axes are the trailing ones and room is left, each such level
                                     is also stretched to the unit's width there (slices x level height x width
                                     floats) so the composition interpolates whole rows                  */
    int64_t levels_scratch_floats;
    skr_pyramid_level levels[SKR_MAX_LEVELS];
} skr_pyramid;

/*
 * Pyramid (noise.py:146-207): out = (base + sum_l weight_l * upsample(level_l)) / std, bilinear/linear
 * upsampling with align_corners=False semantics, unbiased std over the whole tensor.  `moments` = device
 * double[SKR_MOMENTS_DOUBLES], zeroed.  Without `levels_scratch`, two kernels either way:
 *   - without `scratch`: moments pass, then a second pass that regenerates, normalises and writes (nothing
 *     N-sized besides `out` exists; every interpolation corner is a Philox draw: compute-bound);
 *   - with `scratch`: one composition pass writes the unnormalised field to scratch and accumulates the moments,
 *     then a scale pass writes scratch / std to `out`.  With `levels_scratch` as well (and a last extent that is a
 *     multiple of 4) this is the fast path, three launches: the coarse level grids are drawn into levels_scratch, the
 *     composition draws the base and the unit-sized levels in registers and interpolates the coarse grids, the scale
 *     pass normalises.  Buffers supplied for the base / levels are read instead of drawn (skr_noise_fill with the
 *     level's stream gives the same values the in-kernel draw would).
 * `moments`: SKR_MOMENTS_DOUBLES doubles, zeroed.
 */
int skr_noise_pyramid(void* out, int32_t dtype, const skr_pyramid* desc, double* moments, void* cuda_stream);

/*
 * Colored (noise.py:285-335,379-394): multiply an rfftn half-spectrum in place by
 * clamp(normalised radial frequency, 0.5 / max(mean(dims), 4)) ** (-exponent / 2).
 * `spectrum`: complex64 (complex_dtype SKR_F32) or complex128 (SKR_F64), shape dims[0..n-2] x (dims[n-1]/2+1).
 */
int skr_colored_shape(void* spectrum, int32_t complex_dtype, const int64_t* dims, int32_t ndim, double exponent,
                      void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* SKRAMPLE_B200_H */
