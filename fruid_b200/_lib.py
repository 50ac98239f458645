"""ctypes binding of the C ABI declared in include/fruid_b200.h.

There is no CPU fallback: if libfruid_b200.so is missing this module raises at import
time, and every call raises FruidError when the CUDA path fails.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfruid_b200.so")
# The ordering tests run the same Python layer over the FRUID_DEBUG twin (tail delays, poisoned scratch):
# FRUID_B200_LIB=debug selects fruid_b200/lib/libfruid_b200_dbg.so.  Nothing else may be selected: there is
# one product library and no fallback.
if os.environ.get("FRUID_B200_LIB", "") == "debug":
    LIB_PATH = os.path.join(_HERE, "lib", "libfruid_b200_dbg.so")

OK = 0
ERR_BAD_SIZE, ERR_SIZE_MISMATCH, ERR_ODD_ITERATIONS, ERR_CUDA, ERR_NO_PEER, ERR_OOM, ERR_BAD_ARG = range(-1, -8, -1)
POSITIVE, NEGX, NEGY = 0, 1, 2
U, V, U_PREV, V_PREV = 0, 1, 2, 3
DENS, DENS_PREV = 0, 1


class FruidError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fruid_b200 error {code}: {msg}")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -m fruid_b200.build` (nvcc, sm_100a). "
        "fruid_b200 has no CPU fallback.")

lib = ctypes.CDLL(LIB_PATH)

_sz, _f, _i = ctypes.c_size_t, ctypes.c_float, ctypes.c_int
_fp = ctypes.POINTER(ctypes.c_float)


class _vp(ctypes.c_void_p):
    """void* parameter that also takes a numpy array DIRECTLY: `lib.fruid_fluid_upload(h, f, array)`.

    Passing the array itself (rather than `ptr(array)`, an integer) keeps it referenced by the call's argument
    tuple for the duration of the call.  `ptr(np.ascontiguousarray(x))` on a temporary does not: the temporary dies
    when ptr() returns and the library is handed the address of freed memory -- the cause of the "intermittent
    mismatch" that round 1 chased in the kernels."""

    @classmethod
    def from_param(cls, obj):
        if np is not None and isinstance(obj, np.ndarray):  # (np is None while the interpreter shuts down)
            if obj.dtype != np.float32 or not obj.flags.c_contiguous:
                raise TypeError("expected a C-contiguous float32 array")
            return ctypes.c_void_p(obj.ctypes.data)
        if obj is None or isinstance(obj, int):
            return ctypes.c_void_p(obj)
        return obj  # ctypes instances: handles, arrays of handles, byref(...)

# name -> argtypes; every function returns int unless listed in _RESTYPES.
SIGNATURES = {
    "fruid_last_error": [],
    "fruid_kernel_launch_count": [],
    "fruid_build_info": [],
    "fruid_set_tile_config": [_i, _i, _i],
    "fruid_debug_tile_plan": [_sz, _sz, _i, _i, _vp],
    "fruid_set_launch_overlap": [_i],
    "fruid_set_advect_tile": [_i],
    "fruid_debug_set_schedule": [_i],
    "fruid_debug_set_tail_delay": [_i],
    "fruid_debug_recip_mode": [_f, _vp, _vp],
    "fruid_set_graphs_enabled": [_i],
    "fruid_host_register": [_vp, _sz],
    "fruid_host_unregister": [_vp],
    "fruid_profile_enable": [_i],
    "fruid_profile_report": [_vp, _sz, ctypes.POINTER(_sz)],
    "fruid_fluid_stream": [_vp],
    "fruid_density_stream": [_vp],
    "fruid_fluid_create": [_sz, _sz, ctypes.POINTER(_vp)],
    "fruid_fluid_create_slab": [_sz, _sz, _i, _i, ctypes.POINTER(_vp)],
    "fruid_fluid_ipc_export": [_vp, _vp],
    "fruid_fluid_ipc_attach": [_vp, _i, _vp],
    "fruid_fluid_slab_rows": [_vp, ctypes.POINTER(_sz), ctypes.POINTER(_sz)],
    "fruid_fluid_debug_exchange": [_vp, _i, _i],
    "fruid_fluid_debug_barrier": [_vp, _i],
    "fruid_density_create_slab": [_sz, _sz, _i, _i, ctypes.POINTER(_vp)],
    "fruid_density_ipc_export": [_vp, _vp],
    "fruid_density_ipc_attach": [_vp, _i, _vp],
    "fruid_fluid_destroy": [_vp],
    "fruid_fluid_width": [_vp],
    "fruid_fluid_height": [_vp],
    "fruid_fluid_upload": [_vp, _i, _vp],
    "fruid_fluid_download": [_vp, _i, _vp],
    "fruid_fluid_upload_async": [_vp, _i, _vp],
    "fruid_fluid_download_async": [_vp, _i, _vp],
    "fruid_density_upload_async": [_vp, _i, _vp],
    "fruid_density_download_async": [_vp, _i, _vp],
    "fruid_fluid_inject": [_vp, _i, _sz, _sz, _f],
    "fruid_fluid_fill": [_vp, _i, _f],
    "fruid_fluid_set_iterations": [_vp, _i],
    "fruid_fluid_set_fused_sweeps": [_vp, _i],
    "fruid_fluid_step": [_vp, _f, _f],
    "fruid_fluid_step_n": [_vp, _f, _f, _i],
    "fruid_fluid_sync": [_vp],
    "fruid_density_create": [_sz, _sz, ctypes.POINTER(_vp)],
    "fruid_density_destroy": [_vp],
    "fruid_density_upload": [_vp, _i, _vp],
    "fruid_density_download": [_vp, _i, _vp],
    "fruid_density_inject": [_vp, _i, _sz, _sz, _f],
    "fruid_density_fill": [_vp, _i, _f],
    "fruid_density_set_iterations": [_vp, _i],
    "fruid_density_set_fused_sweeps": [_vp, _i],
    "fruid_density_step_dev": [_vp, _vp, _f, _f],
    "fruid_density_step_dev_batch": [_vp, _i, _vp, _f, _f],
    "fruid_density_step_host": [_vp, _vp, _vp, _f, _f],
    "fruid_fluid_bind_mirror": [_vp, _i, _vp],
    "fruid_host_velocity_uploads": [],
    "fruid_density_sync": [_vp],
    "fruid_draw_density": [_vp, _vp, _vp, _vp, _f, _vp],
    "fruid_density_colors": [_vp, _vp, _vp, _vp, _vp],
    "fruid_draw_velocity_lines": [_vp, _f, _vp],
    "fruid_op_add_source": [_vp, _vp, _sz, _sz, _f],
    "fruid_op_set_bnd": [_i, _vp, _sz, _sz],
    "fruid_op_lin_solve": [_i, _vp, _vp, _vp, _sz, _sz, _f, _f, _i, _i],
    "fruid_op_diffuse": [_i, _vp, _vp, _vp, _sz, _sz, _f, _f, _i, _i],
    "fruid_op_advect": [_i, _vp, _vp, _vp, _vp, _sz, _sz, _f],
    "fruid_op_project": [_vp, _vp, _vp, _vp, _vp, _sz, _sz, _i, _i],
    "fruid_dev_lin_solve": [_i, _vp, _vp, _vp, _sz, _sz, _sz, _f, _f, _i, _i, _vp],
    "fruid_dev_advect": [_i, _vp, _vp, _vp, _vp, _sz, _sz, _sz, _f, _vp],
}
_RESTYPES = {
    "fruid_last_error": ctypes.c_char_p,
    "fruid_build_info": ctypes.c_char_p,
    "fruid_kernel_launch_count": ctypes.c_uint64,
    "fruid_host_velocity_uploads": ctypes.c_uint64,
    "fruid_fluid_stream": ctypes.c_void_p,
    "fruid_density_stream": ctypes.c_void_p,
    "fruid_fluid_width": _sz,
    "fruid_fluid_height": _sz,
}
for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library out of sync
    _fn.argtypes = _args
    _fn.restype = _RESTYPES.get(_name, _i)


def check(code: int) -> None:
    if code != OK:
        raise FruidError(code, lib.fruid_last_error().decode("utf-8", "replace"))


def ptr(a: np.ndarray) -> int:
    """Address of a C-contiguous float32 array.  The CALLER must keep `a` referenced until the library call that
    uses the address has returned: never call this on a temporary (pass the array itself instead, see _vp)."""
    if a.dtype != np.float32 or not a.flags.c_contiguous:
        raise TypeError("expected a C-contiguous float32 array")
    return a.ctypes.data


def launch_count() -> int:
    return int(lib.fruid_kernel_launch_count())


class KernelStat(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 48), ("launches", ctypes.c_uint64), ("total_ms", ctypes.c_double)]


def profile_enable(on: bool) -> None:
    check(lib.fruid_profile_enable(1 if on else 0))


def profile_report() -> dict:
    """{kernel name: (launches, total_ms)} since the last report (synchronises the device)."""
    buf = (KernelStat * 64)()
    n = _sz(0)
    check(lib.fruid_profile_report(ctypes.cast(buf, _vp), 64, ctypes.byref(n)))
    return {buf[i].name.decode(): (int(buf[i].launches), float(buf[i].total_ms)) for i in range(n.value)}
