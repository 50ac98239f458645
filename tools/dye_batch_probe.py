"""Scene step with D dye fields (main.rs:114-118): one fruid_density_step_dev per dye against
one fruid_density_step_dev_batch.  Wall-clock per frame (host enqueue included), device-resident.

    python tools/dye_batch_probe.py [size=250] [dyes=4] [frames=300]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import fruid_b200 as fb  # noqa: E402
from fruid_b200 import _lib  # noqa: E402

L = _lib.lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250
D = int(sys.argv[2]) if len(sys.argv) > 2 else 4
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 300
sc = bench.make_scene(n, n)


def setup():
    sim = fb.FluidSim(n, n)
    dyes = [fb.DensitySim(n, n) for _ in range(D)]
    for name, field in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
        _lib.check(L.fruid_fluid_upload(sim._h, field, _lib.ptr(sc[name])))
    for d in dyes:
        for name, field in (("dens", 0), ("dens_prev", 1)):
            _lib.check(L.fruid_density_upload(d._h, field, _lib.ptr(sc[name])))
    return sim, dyes


def run(tag, batched):
    sim, dyes = setup()
    arr = (_lib._vp * D)(*[d._h for d in dyes])

    def frame():
        _lib.check(L.fruid_fluid_step(sim._h, 0.01, 0.0))
        if batched:
            _lib.check(L.fruid_density_step_dev_batch(arr, D, sim._h, 0.01, 0.0))
        else:
            for d in dyes:
                _lib.check(L.fruid_density_step_dev(d._h, sim._h, 0.01, 0.0))

    for _ in range(5):
        frame()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(frames):
        frame()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / frames
    print(f"n={n} dyes={D} {tag:10s} {dt * 1e6:9.1f} us/frame  {n * n / dt / 1e9:7.3f} G cell-steps/s", flush=True)


for rep in range(2):
    run("per-dye", False)
    run("batched", True)
