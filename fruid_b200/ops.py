"""Per-operator entry points (parity tests, profiling): the private fns of src/lib.rs
run through the same CUDA kernels the step uses.  Arrays are (h, w) float32, [y, x]."""
from __future__ import annotations

import numpy as np

from ._lib import check, lib, ptr, POSITIVE, NEGX, NEGY  # noqa: F401


def _wh(a: np.ndarray):
    if a.ndim != 2:
        raise TypeError("expected a 2-D (h, w) float32 array")
    return a.shape[1], a.shape[0]


def add_source(x, s, dt):  # lib.rs:5
    w, h = _wh(x)
    check(lib.fruid_op_add_source(ptr(x), ptr(s), w, h, dt))


def set_bnd(b, x):  # lib.rs:27
    w, h = _wh(x)
    check(lib.fruid_op_set_bnd(b, ptr(x), w, h))


def lin_solve(b, x, x0, scratch, a, c, iters=20, fused=0):  # lib.rs:46
    w, h = _wh(x)
    check(lib.fruid_op_lin_solve(b, ptr(x), ptr(x0), ptr(scratch), w, h, a, c, iters, fused))


def diffuse(b, x, x0, scratch, diff, dt, iters=20, fused=0):  # lib.rs:73
    w, h = _wh(x)
    check(lib.fruid_op_diffuse(b, ptr(x), ptr(x0), ptr(scratch), w, h, diff, dt, iters, fused))


def advect(b, d, d0, u, v, dt):  # lib.rs:84
    w, h = _wh(d)
    check(lib.fruid_op_advect(b, ptr(d), ptr(d0), ptr(u), ptr(v), w, h, dt))


def project(u, v, p, div, scratch, iters=20, fused=0):  # lib.rs:146
    w, h = _wh(u)
    check(lib.fruid_op_project(ptr(u), ptr(v), ptr(p), ptr(div), ptr(scratch), w, h, iters, fused))
