"""Scene-step latency on small grids over tile shapes and fused depths.
    python tools/small_grid_cfg_probe.py 250 512 1024
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, fruid_b200 as fb
from fruid_b200 import _lib
L = _lib.lib
CFGS = [(8, 16), (4, 16), (4, 24), (2, 32), (2, 16), (4, 8), (3, 16)]
if os.environ.get("CFGS"):  # e.g. CFGS=8x16,4x32
    CFGS = [tuple(int(v) for v in c.split("x")) for c in os.environ["CFGS"].split(",")]
REPS = int(os.environ.get("REPS", "200"))
for n in [int(a) for a in sys.argv[1:]] or [250]:
    sc = bench.make_scene(n, n)
    for (R, NW) in CFGS:
        for T in (10, 20):
            if L.fruid_set_tile_config(R, NW, 10) != 0:
                continue
            sim, den = fb.FluidSim(n, n), fb.DensitySim(n, n)
            if L.fruid_fluid_set_fused_sweeps(sim._h, T) != 0 or L.fruid_density_set_fused_sweeps(den._h, T) != 0:
                continue
            for name, field in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
                _lib.check(L.fruid_fluid_upload(sim._h, field, _lib.ptr(sc[name])))
            for name, field in (("dens", 0), ("dens_prev", 1)):
                _lib.check(L.fruid_density_upload(den._h, field, _lib.ptr(sc[name])))
            st = torch.cuda.ExternalStream(int(L.fruid_fluid_stream(sim._h)))
            def both():
                rc = L.fruid_fluid_step(sim._h, 0.01, 0.0) or L.fruid_density_step_dev(den._h, sim._h, 0.01, 0.0)
                return rc
            if both() != 0:
                print(f"n={n} R={R} NW={NW} T={T}: unsupported"); continue
            for _ in range(5): both()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(REPS): both()
            e1.record(st); torch.cuda.synchronize()
            print(f"n={n} R={R} NW={NW} T={T}: {e0.elapsed_time(e1) / REPS * 1e3:8.1f} us/scene step", flush=True)
