"""ctypes loader for the CPU oracle (oracle/fruid_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from fruid_b200/.
PARITY UNPINNED: see the header of fruid_oracle.c.

All arrays are C-contiguous float32 numpy arrays of shape (h, w): element
[y, x] lives at flat index x + y*w, the Array2D layout of the reference
(src/lib.rs:1; idek_basics::Array2D).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "build")

POSITIVE, NEGX, NEGY = 0, 1, 2
FAITHFUL, FAST = 0, 1

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_sz = ctypes.c_size_t
_f = ctypes.c_float
_i = ctypes.c_int


def build(force: bool = False) -> None:
    """Compile the oracle with the committed Makefile (gcc only)."""
    need = force or not os.path.exists(os.path.join(_BUILD, "libfruid_oracle.so"))
    src = os.path.join(_HERE, "fruid_oracle.c")
    so = os.path.join(_BUILD, "libfruid_oracle.so")
    if not need and os.path.getmtime(src) > os.path.getmtime(so):
        need = True
    if need:
        subprocess.run(["make", "-C", _HERE, "-B", "build/libfruid_oracle.so"], check=True,
                       capture_output=True)
        # OpenMP variant is optional (used only to generate big goldens quickly).
        subprocess.run(["make", "-C", _HERE, "-B", "build/libfruid_oracle_omp.so"], check=False,
                       capture_output=True)


def _bind(lib):
    lib.fruid_oracle_add_source.argtypes = [_f32p, _f32p, _sz, _sz, _f]
    lib.fruid_oracle_set_bnd.argtypes = [_i, _f32p, _sz, _sz]
    lib.fruid_oracle_lin_solve.argtypes = [_i, _f32p, _f32p, _f32p, _sz, _sz, _f, _f, _i, _i]
    lib.fruid_oracle_diffuse.argtypes = [_i, _f32p, _f32p, _f32p, _sz, _sz, _f, _f, _i, _i]
    lib.fruid_oracle_advect.argtypes = [_i, _f32p, _f32p, _f32p, _f32p, _sz, _sz, _f, _i]
    lib.fruid_oracle_diff_acc.argtypes = [_f32p, _f32p, _sz, _sz, _f, _i, _i]
    lib.fruid_oracle_project.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, _sz, _sz, _i, _i]
    lib.fruid_oracle_dens_step.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, _sz, _sz, _f, _f, _i, _i]
    lib.fruid_oracle_vel_step.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, _sz, _sz, _f, _f, _i, _i]
    lib.fruid_oracle_draw_density.argtypes = [_f32p, _f32p, _f32p, _f32p, _sz, _sz, _f, _f32p]
    lib.fruid_oracle_draw_density.restype = None
    lib.fruid_oracle_draw_velocity_lines.argtypes = [_f32p, _f32p, _sz, _sz, _f, _f32p]
    lib.fruid_oracle_draw_velocity_lines.restype = None
    lib.fruid_oracle_has_openmp.restype = _i
    for name in ("add_source", "set_bnd", "lin_solve", "diffuse", "advect", "diff_acc", "project",
                 "dens_step", "vel_step"):
        getattr(lib, "fruid_oracle_" + name).restype = None
    return lib


_libs = {}


def lib(openmp: bool = False):
    """Load (building if needed) the serial or the OpenMP oracle library."""
    key = bool(openmp)
    if key not in _libs:
        name = "libfruid_oracle_omp.so" if openmp else "libfruid_oracle.so"
        path = os.path.join(_BUILD, name)
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            if openmp:  # fall back to the serial build; results are identical
                return lib(False)
            raise RuntimeError("oracle library missing and could not be built: " + path)
        _libs[key] = _bind(ctypes.CDLL(path))
    return _libs[key]


def _wh(a):
    assert a.dtype == np.float32 and a.ndim == 2 and a.flags.c_contiguous
    return a.shape[1], a.shape[0]


class Fluid:
    """Oracle twin of FluidSim (src/lib.rs:247-294): u, v, u_prev, v_prev, scratch."""

    def __init__(self, w, h, iters=20, order=FAST, openmp=True):
        self.w, self.h, self.iters, self.order = w, h, iters, order
        self._lib = lib(openmp and order == FAST)
        z = lambda: np.zeros((h, w), np.float32)
        self.u, self.v, self.u_prev, self.v_prev, self.scratch = z(), z(), z(), z(), z()

    def step(self, dt, visc):  # src/lib.rs:267-277
        self._lib.fruid_oracle_vel_step(self.u, self.v, self.u_prev, self.v_prev, self.scratch,
                                        self.w, self.h, visc, dt, self.iters, self.order)


class Density:
    """Oracle twin of DensitySim (src/lib.rs:210-245): dens, dens_prev, scratch."""

    def __init__(self, w, h, iters=20, order=FAST, openmp=True):
        self.w, self.h, self.iters, self.order = w, h, iters, order
        self._lib = lib(openmp and order == FAST)
        z = lambda: np.zeros((h, w), np.float32)
        self.dens, self.dens_prev, self.scratch = z(), z(), z()

    def step(self, uv, dt, diff):  # src/lib.rs:226-236 (note the argument order)
        u, v = uv
        self._lib.fruid_oracle_dens_step(self.dens, self.dens_prev, u, v, self.scratch, self.w,
                                         self.h, diff, dt, self.iters, self.order)
