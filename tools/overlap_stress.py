"""Stress loop for the intermittent density mismatch: the golden 37x29 viscous scenario, many times
in one process, counting mismatching runs.  Knobs come from the environment (read when the library loads):
    FRUID_B200_DENS_OVERLAP=1 FRUID_B200_LAUNCH_OVERLAP=1|0 python tools/overlap_stress.py [iterations]
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fruid_b200 as fb
from fruid_b200 import _lib
L = _lib.lib
g = np.load(os.path.join(ROOT, "tests", "golden", "steps_37x29_visc.npz"))
w, h, n = int(g["w"]), int(g["h"]), int(sys.argv[1]) if len(sys.argv) > 1 else 400
ins = {k: np.ascontiguousarray(g["in_" + k]) for k in ("u", "v", "u_prev", "v_prev", "dens", "dens_prev")}
big = [fb.FluidSim(700, 500) for _ in range(2)]  # some unrelated work in flight, like a test suite leaves behind
bad, rows = 0, {}
for it in range(n):
    sim, den = fb.FluidSim(w, h), fb.DensitySim(w, h)
    for k, f in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
        _lib.check(L.fruid_fluid_upload(sim._h, f, _lib.ptr(ins[k])))
    for k, f in (("dens", 0), ("dens_prev", 1)):
        _lib.check(L.fruid_density_upload(den._h, f, _lib.ptr(ins[k])))
    if it % 3 == 0:
        _lib.check(L.fruid_fluid_step(big[it % 2]._h, 0.01, 0.0))
    for _ in range(int(g["nsteps"])):
        _lib.check(L.fruid_fluid_step(sim._h, float(g["dt"]), float(g["visc"])))
        _lib.check(L.fruid_density_step_dev(den._h, sim._h, float(g["dt"]), float(g["diff"])))
    a = np.empty((h, w), np.float32)
    _lib.check(L.fruid_density_download(den._h, 0, _lib.ptr(a)))
    m = np.nonzero(a.view(np.uint32) != g["out_dens"].view(np.uint32))[0]
    if m.size:
        bad += 1
        for r in set(m.tolist()):
            rows[r] = rows.get(r, 0) + 1
print(f"overlap={os.environ.get('FRUID_B200_DENS_OVERLAP', '0')} pdl={os.environ.get('FRUID_B200_LAUNCH_OVERLAP', '1')}: "
      f"{bad} mismatching runs of {n}; rows hit: {dict(sorted(rows.items()))}")
