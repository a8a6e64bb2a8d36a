"""Step programs: the single description of a solver step that every executor runs.

A sampler step in skrample is a straight-line, elementwise recipe over a few
latent-sized tensors and a handful of float64 scalars.  Instead of issuing one
device op per ``+ - * /`` (what the reference does through ATen), every
sampler here *emits* a short program for a tiny accumulator machine:

    registers  X (sample)  P (converted prediction)  B (base)  A (accumulator)
               S (saved)   R (result)                T, U (scratch)

    ops        LOAD / MOV / STORE, model-space CONVersion, ACCumulate,
               UNI (UniP/UniC divided differences), DPM2/DPM3 corrections,
               FWD/BACK (sample*G + pred*D + noise*Z and its inverse),
               BLEND (SPC), AXPBY (add/remove noise), DIVA

Each op is written so that its arithmetic is *exactly* the reference's op order
with every binary operation individually rounded, which is what makes the CUDA
path bit-identical to the reference torch-CPU path for fp32/fp64.

Two executors exist:

* ``native``  - CUDA tensors.  The whole program is one launch of the sm_100a
  interpreter kernel (``csrc/step_kernel.cu``) through the C ABI.  A missing
  shared library is a hard error, never a silent fallback.
* ``generic`` - Python floats, NumPy arrays and CPU tensors, which the
  reference API also accepts (its RK wrappers enumerate schedule points by
  running samplers on plain floats, reference: skrample/diffusers.py:943-963).
  It evaluates the same ops with ordinary Python operators and
  ``math.sumprod`` so scalar results follow the reference too.
"""

from __future__ import annotations

import dataclasses
import math
import os
from collections.abc import Sequence
from typing import TYPE_CHECKING, Any

if TYPE_CHECKING:
    from skrample_b200.common import Point

# ---------------------------------------------------------------------------------------------
# ISA (kept in lock-step with include/skrample_b200.h)

X, P, B, A, S, R, T, U = range(8)
REG_NAMES = "XPBASRTU"

OP_END = 0
OP_LOAD = 1  # reg[a] = (b&1 ? -in : in)
OP_MOV = 2  # reg[a] = reg[b]
OP_STORE = 3  # out = reg[a]
OP_CONV = 4  # P = conv(X, y); y = in (b == 0) or P (b == 1); a = CONV_* flags; c0,c1,c2
OP_ACC0 = 5  # A = 0 + src*c0; src = in (a == 0) or reg[a-1]
OP_ACC = 6  # A = A + src*c0
OP_DIVA = 7  # A = A / c0
OP_UNI = 8  # A = (a ? 0 : A) + ((in - B)/c0)*c1
OP_UNIC = 9  # A = (a ? 0 : A) + (P - B)*c1
OP_ADDB = 10  # A = B + (a ? 0 : A)
OP_DPM2 = 11  # A = B + c1*(c0*(B - in))
OP_DPM3A = 12  # T = in; U = c0*(B - in)
OP_DPM3B = 13  # d11 = c0*(T - in); d = U - d11; T = U + c1*d; U = c2*d
OP_DPM3C = 14  # A = (B + c0*T) + c1*U
OP_FWD = 15  # R = ((0 + X*c0) + reg[a]*c1) [+ in*c2 if b&1]
OP_BACK = 16  # P = ((R - X*c0) [- in*c2 if b&1]) / c1
OP_BLEND = 17  # X = S*c0 + R*c1 ; a == 1: signed-power blend with c2 = power, c3 = 1/power
OP_AXPBY = 18  # a == 0: R = X*c0 + in*c1 ; a == 1: R = (X - in*c0)/c1

CONV_USE_X = 1
CONV_MUL_X = 2
CONV_MUL_Y = 4
CONV_DIV = 8

MAX_OPS = 64
MAX_INPUTS = 32
MAX_OUTPUTS = 8
MAX_PHILOX_ITEMS = 256


@dataclasses.dataclass(frozen=True, slots=True)
class Op:
    code: int
    a: int = 0
    b: int = 0
    src: int = -1  # index into Program.inputs
    dst: int = -1  # index into Program.outputs
    c: tuple[float, ...] = ()


@dataclasses.dataclass(frozen=True, slots=True)
class ConvSpec:
    """One model-space conversion ``((X*c0 | X) - (y*c1 | y)) [/ c2]`` or ``y*c1`` / ``y/c2``.

    ``None`` stands for the identity.  The flags say which of the (individually
    rounded) operations take part, so the op order of
    reference: skrample/sampling/models.py:92-212 is reproduced exactly.
    """

    flags: int
    c0: float = 0.0
    c1: float = 0.0
    c2: float = 0.0


class Program:
    "Builder for one fused step; inputs are de-duplicated by object identity."

    __slots__ = ("inputs", "ops", "outputs", "philox")

    def __init__(self) -> None:
        self.ops: list[Op] = []
        self.inputs: list[Any] = []
        self.outputs: list[Any] = []  # per output: requested dtype (or None = dtype of first input)
        self.philox: list[Any] = []  # lazy noise draws generated inside the kernel (PhiloxDraw)

    def draw(self, value: Any) -> int:
        for index, known in enumerate(self.philox):
            if known is value:
                return index
        self.philox.append(value)
        return len(self.philox) - 1

    # -- operands
    def input(self, value: Any) -> int:
        for index, known in enumerate(self.inputs):
            if known is value:
                return index
        self.inputs.append(value)
        return len(self.inputs) - 1

    def output(self, dtype: Any = None) -> int:
        self.outputs.append(dtype)
        return len(self.outputs) - 1

    # -- emitters
    def load(self, reg: int, value: Any, negate: bool = False) -> None:
        self.ops.append(Op(OP_LOAD, reg, int(negate), src=self.input(value)))

    def mov(self, dst: int, src: int) -> None:
        if dst != src:
            self.ops.append(Op(OP_MOV, dst, src))

    def store(self, reg: int, dtype: Any = None) -> int:
        slot = self.output(dtype)
        self.ops.append(Op(OP_STORE, reg, dst=slot))
        return slot

    def conv(self, spec: ConvSpec | None, value: Any = None, negate: bool = False) -> None:
        "P = conv(X, value) - or conv(X, P) when ``value`` is None."
        if value is not None:
            if spec is None:
                self.load(P, value, negate)
                return
            if negate:  # negation is folded into the load, then converted from P
                self.load(P, value, True)
                self.ops.append(Op(OP_CONV, spec.flags, 1, c=(spec.c0, spec.c1, spec.c2)))
                return
            self.ops.append(Op(OP_CONV, spec.flags, 0, src=self.input(value), c=(spec.c0, spec.c1, spec.c2)))
        elif spec is not None:
            self.ops.append(Op(OP_CONV, spec.flags, 1, c=(spec.c0, spec.c1, spec.c2)))

    def acc(self, coefficient: float, value: Any = None, reg: int = P, first: bool = False) -> None:
        "A (+)= source*coefficient; source is the tensor ``value`` or register ``reg``."
        code = OP_ACC0 if first else OP_ACC
        if value is not None:
            self.ops.append(Op(code, 0, src=self.input(value), c=(coefficient,)))
        else:
            self.ops.append(Op(code, reg + 1, c=(coefficient,)))

    def diva(self, divisor: float) -> None:
        self.ops.append(Op(OP_DIVA, c=(divisor,)))

    def uni(self, value: Any, rk: float, rho: float, first: bool) -> None:
        self.ops.append(Op(OP_UNI, int(first), src=self.input(value), c=(rk, rho)))

    def unic(self, rho: float, first: bool) -> None:
        self.ops.append(Op(OP_UNIC, int(first), c=(0.0, rho)))

    def addb(self, empty: bool = False) -> None:
        self.ops.append(Op(OP_ADDB, int(empty)))

    def dpm2(self, value: Any, inv_r: float, half: float = 0.5) -> None:
        self.ops.append(Op(OP_DPM2, src=self.input(value), c=(inv_r, half)))

    def dpm3(self, prev1: Any, prev2: Any, inv_r: float, inv_r2: float, mix: float, inv_sum: float, c1: float, c2: float) -> None:
        self.ops.append(Op(OP_DPM3A, src=self.input(prev1), c=(inv_r,)))
        self.ops.append(Op(OP_DPM3B, src=self.input(prev2), c=(inv_r2, mix, inv_sum)))
        self.ops.append(Op(OP_DPM3C, c=(c1, c2)))

    def fwd(self, gamma: float, delta: float, pred: int = A, noise: Any = None, zeta: float = 0.0) -> None:
        if noise is not None and is_lazy_noise(noise):
            if (len(self.philox) >= 2 and not any(known is noise for known in self.philox)) or len(noise.seeds) > MAX_PHILOX_ITEMS:
                noise = noise.materialize()  # the kernel draws at most two noise tensors itself, of at most 256 items each
            else:
                self.ops.append(Op(OP_FWD, pred, 2, src=self.draw(noise), c=(gamma, delta, zeta)))
                return
        if noise is not None:
            self.ops.append(Op(OP_FWD, pred, 1, src=self.input(noise), c=(gamma, delta, zeta)))
        else:
            self.ops.append(Op(OP_FWD, pred, 0, c=(gamma, delta, 0.0)))

    def back(self, gamma: float, delta: float, noise: Any = None, zeta: float = 0.0) -> None:
        if noise is not None:
            self.ops.append(Op(OP_BACK, 0, 1, src=self.input(noise), c=(gamma, delta, zeta)))
        else:
            self.ops.append(Op(OP_BACK, 0, 0, c=(gamma, delta, 0.0)))

    def blend(self, p: float, c: float, power: float = 1.0) -> None:
        if abs(power - 1) > 1e-8:
            self.ops.append(Op(OP_BLEND, 1, c=(p, c, power, 1 / power)))
        else:
            self.ops.append(Op(OP_BLEND, 0, c=(p, c)))

    def axpby(self, value: Any, c0: float, c1: float, remove: bool = False) -> None:
        self.ops.append(Op(OP_AXPBY, int(remove), src=self.input(value), c=(c0, c1)))

    # -- in-kernel noise or a noise tensor?
    SETTLE_ELEMENTS = 1 << 20
    "Below this many elements a step is launch-bound and an in-kernel draw is free."

    def settle_noise(self) -> None:
        """Decide where this step's lazy noise draws (Philox keys) turn into values.  Inside the step kernel removes
        the noise tensor's write and its reads, at ~35 instructions per drawn element.  Steps that stream few
        instructions per byte (Euler, DPM, Adams: bound by HBM) hide that under their loads at any size; the
        divided-difference steps (UniP / UniPC: bound by instruction issue on large latents) would be slowed by
        exactly the time the fill kernel takes, so above ``SETTLE_ELEMENTS`` they read a filled tensor instead."""
        if not self.philox:
            return
        heavy = sum(1 for op in self.ops if op.code in (OP_UNI, OP_UNIC)) >= 2 or any(op.code == OP_BLEND for op in self.ops)
        if not heavy or self.philox[0].numel <= self.SETTLE_ELEMENTS:
            return
        draws, self.philox = self.philox, []
        for index, op in enumerate(self.ops):
            if op.code == OP_FWD and op.b & 2:
                self.ops[index] = Op(OP_FWD, op.a, 1, src=self.input(draws[op.src].materialize()), c=op.c)

    # -- execution
    def run(self) -> list[Any]:
        "Execute and return one value per ``store``."
        return execute(self)


# ---------------------------------------------------------------------------------------------
# dispatch


def _torch() -> Any:
    import sys

    return sys.modules.get("torch")


def is_lazy_noise(value: Any) -> bool:
    "A noise tensor that only exists as (seeds, streams): skrample_b200.pytorch.noise.PhiloxDraw."
    return getattr(value, "is_lazy_noise", False)


_TENSOR: Any = None


def is_cuda_tensor(value: Any) -> bool:
    global _TENSOR
    if _TENSOR is None:
        torch = _torch()
        if torch is None:
            return False
        _TENSOR = torch.Tensor
    return isinstance(value, _TENSOR) and value.is_cuda


def any_cuda(values: Sequence[Any]) -> bool:
    return any(is_cuda_tensor(v) for v in values)


_FUSABLE_DTYPES: tuple[Any, ...] | None = None


def _fusable(values: Sequence[Any]) -> bool:
    "Same-shape floating CUDA tensors on one device: the contract of the fused kernel."
    global _FUSABLE_DTYPES
    torch = _torch()
    if _FUSABLE_DTYPES is None:
        _FUSABLE_DTYPES = (torch.float32, torch.bfloat16, torch.float16, torch.float64)
    first = values[0]
    if not is_cuda_tensor(first):
        return False
    for v in values:
        if not is_cuda_tensor(v) or v.shape != first.shape or v.device != first.device or v.dtype not in _FUSABLE_DTYPES:
            return False
    return True


def execute(program: Program) -> list[Any]:
    if not program.inputs:
        raise ValueError("step program has no inputs")
    if any_cuda(program.inputs):
        if _fusable(program.inputs) and all(d.numel == program.inputs[0].numel() for d in program.philox):
            from skrample_b200 import native

            return native.launch_program(program)
        if os.environ.get("SKRAMPLE_B200_STRICT"):
            raise RuntimeError(
                "skrample_b200: CUDA inputs with mixed shapes/devices/dtypes cannot run as one fused step"
            )
    return execute_generic(program)


# ---------------------------------------------------------------------------------------------
# generic executor (floats, ndarrays, CPU tensors; also device tensors of irregular shape)


def _spowf(x: Any, f: float) -> Any:
    return abs(x) ** f * (-1 * (x < 0) | 1)


def _apply_conv(flags: int, c: tuple[float, ...], x: Any, y: Any) -> Any:
    c0, c1, c2 = c
    if flags & CONV_USE_X:
        # float-on-the-left products as in the reference (``alpha_t * sample``, ``sigma_t * output``)
        lhs = c0 * x if flags & CONV_MUL_X else x
        rhs = c1 * y if flags & CONV_MUL_Y else y
        value = lhs - rhs
    else:
        value = y * c1 if flags & CONV_MUL_Y else y
    return value / c2 if flags & CONV_DIV else value


def execute_generic(program: Program) -> list[Any]:
    reg: list[Any] = [None] * 8
    results: list[Any] = [None] * len(program.outputs)
    values = program.inputs
    torch = _torch()

    # A pending dot product: the reference builds these with math.sumprod, whose float fast path
    # is extended-precision and whose generic path is the plain left-to-right ((0 + a*b) + c*d)...
    pend_p: list[Any] = []
    pend_q: list[Any] = []
    carry: list[Any] = []  # [A] when a dot product continues an accumulator loaded from memory

    def flush() -> None:
        if pend_p:
            if carry:
                total = carry.pop()
                for p_, q_ in zip(pend_p, pend_q, strict=True):
                    total = total + p_ * q_
                reg[A] = total
            else:
                reg[A] = math.sumprod(pend_p, pend_q)
            pend_p.clear()
            pend_q.clear()

    for op in program.ops:
        code = op.code
        c = op.c
        if code in (OP_ACC0, OP_ACC):
            if code == OP_ACC0:
                pend_p.clear()
                pend_q.clear()
                carry.clear()
            elif not pend_p and not carry:
                carry.append(reg[A])
            pend_p.append(values[op.src] if op.a == 0 else reg[op.a - 1])
            pend_q.append(c[0])
            continue
        if code == OP_UNI:
            if op.a:
                pend_p.clear()
                pend_q.clear()
            pend_p.append(c[1])
            pend_q.append((values[op.src] - reg[B]) / c[0])
            continue
        if code == OP_UNIC:
            if op.a:
                pend_p.clear()
                pend_q.clear()
            pend_p.append(c[1])
            pend_q.append(reg[P] - reg[B])
            continue
        flush()

        if code == OP_LOAD:
            v = values[op.src]
            reg[op.a] = -v if op.b & 1 else v
        elif code == OP_MOV:
            reg[op.a] = reg[op.b]
        elif code == OP_STORE:
            out = reg[op.a]
            want = program.outputs[op.dst]
            if not (want is None or isinstance(want, str)) and torch is not None and isinstance(out, torch.Tensor) and out.dtype != want:
                out = out.to(want)
            results[op.dst] = out
        elif code == OP_CONV:
            reg[P] = _apply_conv(op.a, c, reg[X], values[op.src] if op.b == 0 else reg[P])
        elif code == OP_DIVA:
            reg[A] = reg[A] / c[0]
        elif code == OP_ADDB:
            reg[A] = reg[B] + (0 if op.a else reg[A])
        elif code == OP_DPM2:
            reg[A] = reg[B] + c[1] * (c[0] * (reg[B] - values[op.src]))
        elif code == OP_DPM3A:
            reg[T] = values[op.src]
            reg[U] = c[0] * (reg[B] - reg[T])
        elif code == OP_DPM3B:
            d11 = c[0] * (reg[T] - values[op.src])
            d10 = reg[U]
            reg[T] = d10 + c[1] * (d10 - d11)
            reg[U] = c[2] * (d10 - d11)
        elif code == OP_DPM3C:
            reg[A] = reg[B] + c[0] * reg[T] + c[1] * reg[U]
        elif code == OP_FWD:
            if op.b & 2:
                drawn = program.philox[op.src].materialize()
                reg[R] = math.sumprod((reg[X], reg[op.a], drawn), (c[0], c[1], c[2]))
            elif op.b & 1:
                reg[R] = math.sumprod((reg[X], reg[op.a], values[op.src]), (c[0], c[1], c[2]))
            else:
                reg[R] = math.sumprod((reg[X], reg[op.a]), (c[0], c[1]))
        elif code == OP_BACK:
            if op.b & 1:
                reg[P] = (reg[R] - reg[X] * c[0] - values[op.src] * c[2]) / c[1]
            else:
                reg[P] = (reg[R] - reg[X] * c[0]) / c[1]
        elif code == OP_BLEND:
            if op.a:
                reg[X] = _spowf(_spowf(reg[S], c[2]) * c[0] + _spowf(reg[R], c[2]) * c[1], c[3])
            else:
                reg[X] = reg[S] * c[0] + reg[R] * c[1]
        elif code == OP_AXPBY:
            if op.a:
                reg[R] = (reg[X] - values[op.src] * c[0]) / c[1]
            else:
                reg[R] = reg[X] * c[0] + values[op.src] * c[1]
        else:
            raise ValueError(f"unknown step-program op {code}")
    return results


# ---------------------------------------------------------------------------------------------
# Point.add_noise / remove_noise (reference: skrample/common.py:32-40)


def point_add_noise(point: "Point", sample: Any, noise: Any) -> Any:
    if is_cuda_tensor(sample) and is_cuda_tensor(noise) and _fusable((sample, noise)):
        prog = Program()
        prog.load(X, sample)
        prog.axpby(noise, point.alpha, point.sigma)
        prog.store(R)
        return prog.run()[0]
    return sample * point.alpha + noise * point.sigma


def point_remove_noise(point: "Point", sample: Any, noise: Any) -> Any:
    if is_cuda_tensor(sample) and is_cuda_tensor(noise) and _fusable((sample, noise)):
        prog = Program()
        prog.load(X, sample)
        prog.axpby(noise, point.sigma, point.alpha, remove=True)
        prog.store(R)
        return prog.run()[0]
    scaled = noise * point.sigma
    try:
        return (sample - scaled) / point.alpha
    except ZeroDivisionError:
        return scaled
