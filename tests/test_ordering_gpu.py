"""Ordering regression test: every launch schedule of the library -- graph replay or
direct launches, programmatic dependent launch on or off, the density diffuse beside the velocity step or after it,
the u / v diffuse fork -- must give the oracle's bits on the FRUID_DEBUG build, where the lowest warps of every
block stall before their stores and scratch buffers hold NaNs.  A missing dependency between a producer and a
consumer kernel is then a deterministic mismatch.  Runs tools/order_probe.py in a subprocess because the debug
library is selected at import time (FRUID_B200_LIB=debug)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_every_schedule_is_bit_exact_with_delayed_stores(tmp_path):
    out = tmp_path / "order.json"
    env = dict(os.environ, FRUID_B200_LIB="debug")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "order_probe.py"), "--quick", "--assert-safe",
                        "--json", str(out)], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    s = json.load(open(out))["summary"]
    assert "FRUID_DEBUG" in s["build"] and s["delay_ns"] > 0
    assert s["chained_mismatches"] == 0


@pytest.mark.gpu
def test_repeated_steps_without_gc_between_objects(fb, oracle_mod):
    """Handles created and dropped back to back (no gc.collect() in between, unlike the suite's fixture used to do):
    a new simulation may start while the previous one's teardown has only just been issued."""
    import numpy as np
    from conftest import assert_bits
    rng = np.random.default_rng(5)
    w, h = 37, 29
    init = [rng.standard_normal((h, w)).astype(np.float32) for _ in range(6)]
    of, od = oracle_mod.Fluid(w, h), oracle_mod.Density(w, h)
    of.u[:], of.v[:], of.u_prev[:], of.v_prev[:], od.dens[:], od.dens_prev[:] = init
    for _ in range(3):
        of.step(0.05, 2e-4)
        od.step((of.u, of.v), 0.05, 1e-4)
    L = fb._lib.lib
    for rep in range(40):
        sim, den = fb.FluidSim(w, h), fb.DensitySim(w, h)
        for a, f in zip(init[:4], range(4)):
            fb._lib.check(L.fruid_fluid_upload(sim._h, f, fb._lib.ptr(a)))
        for a, f in zip(init[4:], range(2)):
            fb._lib.check(L.fruid_density_upload(den._h, f, fb._lib.ptr(a)))
        for _ in range(3):
            sim.step(0.05, 2e-4)
            den.step(sim.uv(), 0.05, 1e-4)
        assert_bits(den.density().numpy(), od.dens, f"dens (repeat {rep})")
        assert_bits(den.density_mut().numpy(), od.dens_prev, f"dens_prev (repeat {rep})")
        assert_bits(sim.uv()[0].numpy(), of.u, f"u (repeat {rep})")
