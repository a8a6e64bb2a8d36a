"""Step-plan cache: skip re-deriving a step program when the same step is taken again.

Emitting a step program costs tens to hundreds of microseconds of Python (schedule points, lambda ratios, the
UniPC linear solve, op objects, ctypes packing) - far more than the kernel it produces at SDXL sizes.  The
program of a step depends only on (sampler, model, schedule, step, steps of the history entries, which optional
tensors are present); the tensors themselves only supply pointers.  So the first time a step is taken its
packed program and the *roles* of its inputs (sample / prediction / noise / field j of history entry k) are
remembered, and later calls with the same key just bind pointers and launch.

The cache is per thread (the packed struct is patched in place), keyed on object identity of the frozen
sampler / model / schedule dataclasses (verified with ``is`` against strong references held by the plan) and holds
at most a few hundred plans.  Only device tensors take this path.
"""

from __future__ import annotations

import threading
from typing import Any

SAMPLE, PREDICTION, NOISE, XHAT = 0, 1, 2, 3
MATERIALISED = 9  # role (MATERIALISED, inner role): the tensor of a lazy noise draw, written by its fill kernel on demand
_FIELDS = ("sample", "prediction", "noise")
_MAX_PLANS = 512
XHAT_ATTR = "_skr_xhat"


class Plan:
    __slots__ = ("anchors", "compiled", "result", "roles")

    def __init__(self, anchors: tuple, compiled: Any, roles: tuple, result: tuple) -> None:
        self.anchors = anchors
        self.compiled = compiled
        self.roles = roles
        self.result = result


class _Local(threading.local):
    def __init__(self) -> None:
        self.plans: dict[tuple, Plan] = {}
        self.hits = 0
        self.misses = 0


_local = _Local()


def stats() -> tuple[int, int]:
    "(hits, misses) of this thread's plan cache."
    return _local.hits, _local.misses


def clear() -> None:
    _local.plans.clear()


def _noise_kind(noise: Any) -> int:
    return 0 if noise is None else (2 if getattr(noise, "is_lazy_noise", False) else 1)


_ENTRY_KEY = "_skr_plan_key"
_KEEP = "_skr_keep"


def _entry_key(entry: Any) -> tuple:
    "What a history entry contributes to the key: its step, the kind of its noise, whether it carries an x-hat cache."
    held = entry.__dict__
    key = held.get(_ENTRY_KEY)
    if key is None:  # entries are frozen: computed once (the x-hat cache is attached before an entry can be history)
        key = held[_ENTRY_KEY] = (entry.step, _noise_kind(entry.noise), XHAT_ATTR in held)
    elif key[2] != (XHAT_ATTR in held):  # an x-hat was converted on demand and remembered since
        key = held[_ENTRY_KEY] = (entry.step, key[1], XHAT_ATTR in held)
    return key


def history_depth(sampler: Any) -> int:
    "``sampler.require_previous``, remembered on the (frozen) sampler object: the property walks the sampler tree."
    keep = sampler.__dict__.get(_KEEP)
    if keep is None:
        keep = sampler.require_previous
        object.__setattr__(sampler, _KEEP, keep)
    return keep


def key_for(sampler: Any, packed: Any, model_transform: Any, schedule: Any, previous: Any, out_dtype: Any) -> tuple:
    count = len(previous)
    tail: tuple = ()
    if count:
        keep = sampler.__dict__.get(_KEEP)
        if keep is None:
            keep = history_depth(sampler)
        if keep:
            tail = tuple([_entry_key(p) for p in (previous[count - keep :] if keep < count else previous)])
    noise = packed.noise
    return (
        id(sampler),
        id(model_transform),
        id(schedule),
        packed.step,
        count,
        0 if noise is None else (2 if getattr(noise, "is_lazy_noise", False) else 1),
        tail,
        out_dtype,
    )


def lookup(key: tuple, sampler: Any, model_transform: Any, schedule: Any) -> Plan | None:
    plan = _local.plans.get(key)
    if plan is None:
        _local.misses += 1
        return None
    a = plan.anchors
    if a[0] is not sampler or a[1] is not model_transform or a[2] is not schedule:
        del _local.plans[key]  # an id was recycled by a different object
        _local.misses += 1
        return None
    _local.hits += 1
    return plan


def roles_of(inputs: list[Any], packed: Any, previous: Any) -> tuple | None:
    "Where every program input (tensor or lazy noise draw) came from, or None when one is not the step's own."
    where: dict[int, tuple] = {}

    def note(value: Any, role: tuple) -> bool:
        lazy = getattr(value, "is_lazy_noise", False)
        if value is None or not (hasattr(value, "data_ptr") or lazy):
            return True
        if id(value) in where:
            return False  # one tensor in two roles: a later call may pass different tensors, do not cache
        where[id(value)] = role
        if lazy and value._tensor is not None:  # a draw the step chose to read from memory (Program.settle_noise)
            where[id(value._tensor)] = (MATERIALISED, role)
        return True

    ok = note(packed.sample, (SAMPLE,)) and note(packed.prediction, (PREDICTION,)) and note(packed.noise, (NOISE,))
    for back in range(1, len(previous) + 1):
        entry = previous[-back]
        for f, name in enumerate(_FIELDS):
            ok = ok and note(getattr(entry, name), (4, back, f))
        held = entry.__dict__.get(XHAT_ATTR)
        if held is not None and held[1] is not entry.prediction:  # a trivial conversion remembers the prediction itself
            ok = ok and note(held[1], (4, back, XHAT, held[0]))
    if not ok:
        return None
    roles = []
    for value in inputs:
        role = where.get(id(value))
        if role is None:
            return None
        roles.append(role)
    return tuple(roles)


def store(key: tuple, anchors: tuple, compiled: Any, roles: tuple, result: tuple) -> None:
    if len(_local.plans) >= _MAX_PLANS:
        _local.plans.clear()
    _local.plans[key] = Plan(anchors, compiled, roles, result)


def bind(plan: Plan, packed: Any, previous: Any) -> list[Any] | None:
    "The input tensors of this call in program order, or None when the plan does not apply after all."
    bound = []
    add = bound.append
    for role in plan.roles:
        kind = role[0]
        if kind == SAMPLE:
            add(packed.sample)
        elif kind == PREDICTION:
            add(packed.prediction)
        elif kind == NOISE:
            add(packed.noise)
        elif kind == MATERIALISED:
            inner = role[1]
            draw = packed.noise if inner[0] == NOISE else previous[-inner[1]].noise
            if not getattr(draw, "is_lazy_noise", False):
                return None
            add(draw.materialize())  # remembered on the draw: the next step's corrector reads the same tensor
        else:
            entry = previous[-role[1]]
            field = role[2]
            if field == XHAT:
                held = entry.__dict__.get(XHAT_ATTR)
                if held is None or (held[0] is not role[3] and held[0] != role[3]):
                    return None
                add(held[1])
            elif field == 0:
                add(entry.sample)
            elif field == 1:
                add(entry.prediction)
            else:
                add(entry.noise)
    return bound
