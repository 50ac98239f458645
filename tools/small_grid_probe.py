import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, fruid_b200 as fb
from fruid_b200 import _lib
L = _lib.lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
if len(sys.argv) > 2:
    L.fruid_set_launch_overlap(int(sys.argv[2]))
    print("launch overlap (PDL):", int(sys.argv[2]))
sc = bench.make_scene(n, n)
def setup():
    sim, den = fb.FluidSim(n, n), fb.DensitySim(n, n)
    for name, field in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
        _lib.check(L.fruid_fluid_upload(sim._h, field, _lib.ptr(sc[name])))
    for name, field in (("dens", 0), ("dens_prev", 1)):
        _lib.check(L.fruid_density_upload(den._h, field, _lib.ptr(sc[name])))
    return sim, den
def timeit(tag, fn, sim, steps=300):
    st = torch.cuda.ExternalStream(int(L.fruid_fluid_stream(sim._h)))
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps): fn()
    e1.record(st); torch.cuda.synchronize()
    print(f"{tag:40s} {e0.elapsed_time(e1)/steps*1e3:8.1f} us/step", flush=True)
for graphs in (1, 0):
    L.fruid_set_graphs_enabled(graphs)
    sim, den = setup()
    timeit(f"n={n} graphs={graphs} fluid only", lambda: _lib.check(L.fruid_fluid_step(sim._h, 0.01, 0.0)), sim)
    def both():
        _lib.check(L.fruid_fluid_step(sim._h, 0.01, 0.0)); _lib.check(L.fruid_density_step_dev(den._h, sim._h, 0.01, 0.0))
    timeit(f"n={n} graphs={graphs} fluid + density", both, sim)
    _lib.profile_enable(True)
    for _ in range(10): both()
    st = _lib.profile_report(); _lib.profile_enable(False)
    print("   per-kernel us:", {k: round(t / nn * 1e3, 1) for k, (nn, t) in sorted(st.items())}, flush=True)
