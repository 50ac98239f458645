"""Does exporting a cudaMalloc allocation over CUDA IPC change how fast local gathers run?"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, fruid_b200 as fb
from fruid_b200 import _lib
L = _lib.lib
w = h = 4096
sc = bench.make_scene(w, h)
def run(tag, export):
    sim, den = fb.FluidSim(w, h), fb.DensitySim(w, h)
    if export:
        buf = (ctypes.c_ubyte * 64)()
        _lib.check(L.fruid_fluid_ipc_export(sim._h, ctypes.cast(buf, ctypes.c_void_p)))
        _lib.check(L.fruid_density_ipc_export(den._h, ctypes.cast(buf, ctypes.c_void_p)))
    for name, field in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
        _lib.check(L.fruid_fluid_upload(sim._h, field, _lib.ptr(sc[name])))
    for name, field in (("dens", 0), ("dens_prev", 1)):
        _lib.check(L.fruid_density_upload(den._h, field, _lib.ptr(sc[name])))
    for _ in range(3):
        _lib.check(L.fruid_fluid_step(sim._h, 0.01, 0.0)); _lib.check(L.fruid_density_step_dev(den._h, sim._h, 0.01, 0.0))
    _lib.profile_enable(True)
    for _ in range(5):
        _lib.check(L.fruid_fluid_step(sim._h, 0.01, 0.0)); _lib.check(L.fruid_density_step_dev(den._h, sim._h, 0.01, 0.0))
    st = _lib.profile_report(); _lib.profile_enable(False)
    print(tag, {k: round(t / 5, 4) for k, (n, t) in sorted(st.items())}, flush=True)
run("plain   ", False)
run("exported", True)
run("plain   ", False)
