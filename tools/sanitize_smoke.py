"""A short tour through every kernel of the library at small sizes (all tile shapes, every division mode, batched dyes,
rendering kernels, fill / inject, asynchronous transfers), each result compared with the oracle.

Written to run under `compute-sanitizer --tool memcheck|racecheck`; that tool is closed on this GPU pool (the launcher
answers "compute-sanitizer is closed on this pool"), so here it is a plain quick parity tour:

    python tools/sanitize_smoke.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import fruid_b200 as fb  # noqa: E402
import oracle  # noqa: E402
from fruid_b200 import _lib  # noqa: E402

L = _lib.lib
bad = 0


def same(a, b, what):
    global bad
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    n = int(np.count_nonzero((a.view(np.uint32) != b.view(np.uint32)) & ~(np.isnan(a) & np.isnan(b))))
    print(f"  {what:48s} {'ok' if n == 0 else f'{n} MISMATCHES'}", flush=True)
    bad += n != 0


def scene_steps(n, m, visc, steps, dyes, pin=None):
    if pin:
        _lib.check(L.fruid_set_tile_config(*pin))
    sc = bench.make_scene(n, m)
    sim, osim = fb.FluidSim(n, m), oracle.Fluid(n, m)
    ds = [fb.DensitySim(n, m) for _ in range(dyes)]
    ods = [oracle.Density(n, m) for _ in range(dyes)]
    for name, field in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
        _lib.check(L.fruid_fluid_upload(sim._h, field, _lib.ptr(sc[name])))
        getattr(osim, name)[:] = sc[name]
    for d, od in zip(ds, ods):
        for name, field in (("dens", 0), ("dens_prev", 1)):
            _lib.check(L.fruid_density_upload(d._h, field, _lib.ptr(sc[name])))
            getattr(od, name)[:] = sc[name]
    for _ in range(steps):
        sim.step(0.01, visc)
        osim.step(0.01, visc)
        fb.DensitySim.step_all(ds, sim.uv(), 0.01, visc)
        for od in ods:
            od.step((osim.u, osim.v), 0.01, visc)
    tag = f"{n}x{m} visc={visc} dyes={dyes} pin={pin}"
    same(sim.uv()[0].numpy(), osim.u, f"u {tag}")
    same(sim.uv()[1].numpy(), osim.v, f"v {tag}")
    same(ds[-1].density().numpy(), ods[-1].dens, f"dens {tag}")
    if pin:
        _lib.check(L.fruid_set_tile_config(0, 0, 10))
    return sim, ds


print("scene steps (direct, captured, replayed):")
scene_steps(5, 2, 0.0, 2, 1)                       # degenerate grid: serial kernels
scene_steps(96, 80, 0.0, 3, 4)                     # 128x32 tiles, batched dyes
scene_steps(250, 250, 1e-4, 3, 2)                  # general divisor (reciprocal division + its check)
scene_steps(520, 300, 0.0, 3, 1)                   # 128x48 / 128x64 tiles
scene_steps(300, 300, 0.0, 3, 1, pin=(8, 16, 10))  # the large-grid shape, two launches per solve
scene_steps(200, 150, 0.0, 2, 1, pin=(4, 24, 20))  # more than 16 warps per CTA
sim, ds = scene_steps(130, 70, 0.0, 2, 4)

print("stand-alone operators:")
rng = np.random.default_rng(3)
w, h = 77, 41
O = oracle.lib(True)
for b in (0, 1, 2):
    x = rng.standard_normal((h, w)).astype(np.float32)
    xo = x.copy()
    fb.ops.set_bnd(b, x)
    O.fruid_oracle_set_bnd(b, xo, w, h)
    same(x, xo, f"set_bnd b={b}")
x, x0, s = (rng.standard_normal((h, w)).astype(np.float32) for _ in range(3))
xo, so = x.copy(), s.copy()
fb.ops.lin_solve(1, x, x0, s, 0.3, 2.2, 20, 1)     # one sweep per launch
O.fruid_oracle_lin_solve(1, xo, x0, so, w, h, 0.3, 2.2, 20, 1)
same(x, xo, "lin_solve fused=1")
a, src = rng.standard_normal((h, w)).astype(np.float32), rng.standard_normal((h, w)).astype(np.float32)
ao = a.copy()
fb.ops.add_source(a, src, 0.1)
O.fruid_oracle_add_source(ao, src, w, h, 0.1)
same(a, ao, "add_source")

print("rendering kernels, fill / inject, asynchronous transfers:")
verts = fb.render.draw_density(*ds, z=0.5)
want = np.zeros(verts.size, np.float32)
O.fruid_oracle_draw_density(*[np.ascontiguousarray(d.density().numpy()) for d in ds], 130, 70, 0.5, want)
same(verts[:, :3], want.reshape(-1, 6)[:, :3], "draw_density positions")
lines = fb.render.draw_velocity_lines(sim, z=0.25)
want = np.zeros(lines.size, np.float32)
O.fruid_oracle_draw_velocity_lines(np.ascontiguousarray(sim.uv()[0].numpy()), np.ascontiguousarray(sim.uv()[1].numpy()), 130, 70,
                                   0.25, want)
same(lines, want.reshape(-1, 6), "draw_velocity_lines")
ds[0].density_mut().fill(0.0)
ds[1].density_mut()[(5, 6)] = 3.0
_lib.check(L.fruid_density_fill(ds[2]._h, 1, 1.5))
sim.step(0.01, 0.0)
fb.DensitySim.step_all(ds, sim.uv(), 0.01, 0.0)
host = np.zeros((70, 130), np.float32)
pinned = np.ones((70, 130), np.float32)
_lib.check(L.fruid_density_upload_async(ds[3]._h, 1, _lib.ptr(pinned)))
_lib.check(L.fruid_density_step_dev(ds[3]._h, sim._h, 0.01, 0.0))
_lib.check(L.fruid_density_download_async(ds[3]._h, 0, _lib.ptr(host)))
_lib.check(L.fruid_density_sync(ds[3]._h))
print("  async round trip finite:", bool(np.isfinite(host).all()))
print("SANITIZE SMOKE", "OK" if bad == 0 else f"FAILED ({bad})")
sys.exit(0 if bad == 0 else 1)
