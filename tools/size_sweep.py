"""Scene-step throughput vs grid size on one GPU (device-resident, CUDA-event timed)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, fruid_b200 as fb
from fruid_b200 import _lib
L = _lib.lib
for n in ([int(a) for a in sys.argv[1:]] or [128, 250, 512, 1024, 2048, 4096, 8192, 16384]):
    sc = bench.make_scene(n, n)
    sim, den = fb.FluidSim(n, n), fb.DensitySim(n, n)
    for name, field in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
        _lib.check(L.fruid_fluid_upload(sim._h, field, _lib.ptr(sc[name])))
    for name, field in (("dens", 0), ("dens_prev", 1)):
        _lib.check(L.fruid_density_upload(den._h, field, _lib.ptr(sc[name])))
    st = torch.cuda.ExternalStream(int(L.fruid_fluid_stream(sim._h)))
    def step():
        _lib.check(L.fruid_fluid_step(sim._h, 0.01, 0.0)); _lib.check(L.fruid_density_step_dev(den._h, sim._h, 0.01, 0.0))
    for _ in range(5): step()
    torch.cuda.synchronize()
    steps = 200 if n <= 1024 else (20 if n <= 4096 else 5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps): step()
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{n:6d}^2: {ms*1e3:9.1f} us/step  {n*n/ms/1e6:8.2f} G cell-steps/s", flush=True)
    del sim, den
