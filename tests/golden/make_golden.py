"""Generates the committed golden fixtures from the CPU oracle (oracle/fruid_oracle.c).

The reference ships no golden vectors and cannot be run here (Rust toolchain absent), so
these fixtures pin the ORACLE (and through it the CUDA path) against regressions; they do
not pin the oracle to the reference ("parity unpinned", see oracle/fruid_oracle.c).

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz, *.json
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from oracle_api import ODensitySim, OFluidSim  # noqa: E402
from fruid_b200_scene_import import Scene  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.float32).tobytes()).hexdigest()


def steps_fixture(w, h, seed, visc, diff, dt, nsteps):
    rng = np.random.default_rng(seed)
    f = oracle.Fluid(w, h, order=oracle.FAITHFUL)
    d = oracle.Density(w, h, order=oracle.FAITHFUL)
    for a in (f.u, f.v, f.u_prev, f.v_prev, d.dens, d.dens_prev):
        a[:] = rng.standard_normal((h, w)).astype(np.float32)
    out = {"w": w, "h": h, "visc": np.float32(visc), "diff": np.float32(diff), "dt": np.float32(dt),
           "nsteps": nsteps}
    for k, a in (("u", f.u), ("v", f.v), ("u_prev", f.u_prev), ("v_prev", f.v_prev), ("dens", d.dens),
                 ("dens_prev", d.dens_prev)):
        out["in_" + k] = a.copy()
    for _ in range(nsteps):
        f.step(dt, visc)
        d.step((f.u, f.v), dt, diff)
    for k, a in (("u", f.u), ("v", f.v), ("u_prev", f.u_prev), ("v_prev", f.v_prev), ("dens", d.dens),
                 ("dens_prev", d.dens_prev)):
        out["out_" + k] = a.copy()
    return out


def scene_fields(s):
    u, v = s.sim.uv()
    p, dv = s.sim.uv_mut()
    fields = {"u": u.numpy(), "v": v.numpy(), "u_prev": p.numpy(), "v_prev": dv.numpy()}
    for i, d in enumerate(s.dyes):
        fields[f"dens{i}"] = d.density().numpy()
        fields[f"dens_prev{i}"] = d.density_mut().numpy()
    return fields


def main():
    oracle.build()
    np.savez_compressed(os.path.join(HERE, "steps_37x29_visc.npz"), **steps_fixture(37, 29, 1234, 2e-4, 1e-4, 0.05, 3))
    np.savez_compressed(os.path.join(HERE, "steps_30x41_inviscid.npz"), **steps_fixture(30, 41, 99, 0.0, 0.0, 0.1, 2))
    # reference scene, small enough to commit whole
    s = Scene(OFluidSim, ODensitySim, 64, 48)
    s.run(6)
    np.savez_compressed(os.path.join(HERE, "scene_64x48_f6.npz"), **scene_fields(s))
    # checksums of the faithful default scene (src/main.rs:36: 250x250) and the 128x128 variant
    sums = {}
    for (w, h, frames) in ((250, 250, 100), (128, 128, 100)):
        s = Scene(OFluidSim, ODensitySim, w, h)
        s.run(frames)
        fl = scene_fields(s)
        sums[f"{w}x{h}_f{frames}"] = {
            "sha256": {k: sha(a) for k, a in fl.items()},
            "probe": {k: float(a[2 * h // 3, 2 * w // 3]) for k, a in fl.items()},  # cf. dbg! at main.rs:54
            "absmax": {k: float(np.abs(a).max()) for k, a in fl.items()},
        }
    # BASELINE.json configs[1]: 1024 x 1024, one dye, the reference's injection law, 1000 frames (and a 200-frame run with
    # visc = diff = 1e-6, SURVEY 8d "variant B"); the oracle's OpenMP / cache-friendly traversal, bit-identical to the
    # faithful one (tests/test_oracle.py)
    for (w, h, frames, visc, diff) in ((1024, 1024, 1000, 0.0, 0.0), (1024, 1024, 200, 1e-6, 1e-6)):
        s = Scene(OFluidSim, ODensitySim, w, h, n_dyes=1)
        s.run(frames, visc=visc, diff=diff)
        fl = scene_fields(s)
        sums[f"{w}x{h}_f{frames}_d1" + ("" if visc == 0.0 else "_visc")] = {
            "sha256": {k: sha(a) for k, a in fl.items()},
            "probe": {k: float(a[2 * h // 3, 2 * w // 3]) for k, a in fl.items()},
            "absmax": {k: float(np.abs(a).max()) for k, a in fl.items()},
        }
    with open(os.path.join(HERE, "scene_checksums.json"), "w") as fh:
        json.dump(sums, fh, indent=1, sort_keys=True)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
