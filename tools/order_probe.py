"""Ordering probe: runs whole scene steps against the oracle under every launch schedule the library has, on the
FRUID_DEBUG build (fruid_b200/lib/libfruid_b200_dbg.so), where the lowest warps of every block stall before they
store and scratch buffers are poisoned with NaNs.  A consumer kernel that starts before its producer has COMPLETED
-- a missing stream, event or programmatic dependency -- then yields a mismatch every time instead of once in a few
hundred runs.

    FRUID_B200_LIB=debug python tools/order_probe.py [--delay-ns 30000] [--assert-safe] [--json out.json]

Schedules: graphs on/off x programmatic launch on/off x density/velocity overlap on/off x u/v fork on/off, each with
the chained launch attribution (this round) and, for reference, the round-1 attribution ("unchained": the programmatic
attribute on every launch, also straight after a graph launch / event wait).  --assert-safe exits 1 if any CHAINED
schedule mismatches; unchained results are reported, not asserted (they document the round-1 race).
"""
import argparse
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--delay-ns", type=int, default=30000)
ap.add_argument("--assert-safe", action="store_true")
ap.add_argument("--json", default="")
ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--quick", action="store_true", help="fewer schedules (the pytest wrapper)")
ap.add_argument("--only", default="", help="run only the cases whose name contains this")
args = ap.parse_args()

import fruid_b200 as fb  # noqa: E402
import oracle  # noqa: E402
from fruid_b200 import _lib  # noqa: E402

L = _lib.lib
info = L.fruid_build_info().decode()
debug_build = "FRUID_DEBUG" in info
GOLD = os.path.join(ROOT, "tests", "golden")


DETAILS = []


def bits_bad(a, b, name=""):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    bad = ~((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)))
    n = int(np.count_nonzero(bad))
    if n and len(DETAILS) < 12:
        ys, xs = np.nonzero(bad)
        with np.errstate(all="ignore"):
            rel = float(np.nanmax(np.abs(a - b)[bad] / np.maximum(np.abs(b[bad]), 1e-30))) if n else 0.0
        DETAILS.append({"field": name, "cells": n, "nan_got": int(np.isnan(a).sum()), "rows": sorted(set(ys.tolist()))[:12],
                        "max_rel": rel, "first": [(int(x), int(y), float(a[y, x]), float(b[y, x])) for y, x in list(zip(ys, xs))[:4]]})
    return n


def case_golden(name):
    g = np.load(os.path.join(GOLD, name))
    w, h = int(g["w"]), int(g["h"])
    sim, den = fb.FluidSim(w, h), fb.DensitySim(w, h)
    # the arrays must stay referenced across the C call: `ptr(np.ascontiguousarray(..))` on a temporary hands the
    # library the address of memory that Python has already freed (the round-1 "intermittent mismatch")
    keep = {k: np.ascontiguousarray(g["in_" + k]) for k in ("u", "v", "u_prev", "v_prev", "dens", "dens_prev")}
    if os.environ.get("FRUID_PROBE_TEMPORARIES") == "1":  # reproduce the round-1 test bug on purpose
        for k, f in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
            _lib.check(L.fruid_fluid_upload(sim._h, f, _lib.ptr(np.ascontiguousarray(g["in_" + k]))))
        for k, f in (("dens", 0), ("dens_prev", 1)):
            _lib.check(L.fruid_density_upload(den._h, f, _lib.ptr(np.ascontiguousarray(g["in_" + k]))))
    else:
        for k, f in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
            _lib.check(L.fruid_fluid_upload(sim._h, f, _lib.ptr(keep[k])))
        for k, f in (("dens", 0), ("dens_prev", 1)):
            _lib.check(L.fruid_density_upload(den._h, f, _lib.ptr(keep[k])))
    for _ in range(int(g["nsteps"]) ):
        sim.step(float(g["dt"]), float(g["visc"]))
        den.step(sim.uv(), float(g["dt"]), float(g["diff"]))
    bad = {}
    for k, f in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
        a = np.empty((h, w), np.float32)
        _lib.check(L.fruid_fluid_download(sim._h, f, _lib.ptr(a)))
        n = bits_bad(a, g["out_" + k], k)
        if n:
            bad[k] = n
    for k, f in (("dens", 0), ("dens_prev", 1)):
        a = np.empty((h, w), np.float32)
        _lib.check(L.fruid_density_download(den._h, f, _lib.ptr(a)))
        n = bits_bad(a, g["out_" + k], k)
        if n:
            bad[k] = n
    return bad


_oracle_cache = {}


def case_random(w, h, steps, visc, diff, ndyes, seed, change_dt=False):
    """`steps` scene steps (1 velocity + ndyes batched dyes) from seeded random fields against the oracle; with
    change_dt the time step alternates, which forces direct launches between graph replays."""
    rng = np.random.default_rng(seed)
    init = [rng.standard_normal((h, w)).astype(np.float32) for _ in range(4 + 2 * ndyes)]
    dts = [0.02 if (change_dt and i % 3 == 2) else 0.01 for i in range(steps)]
    key = (w, h, steps, visc, diff, ndyes, seed, change_dt)
    if key not in _oracle_cache:
        of = oracle.Fluid(w, h)
        ods = [oracle.Density(w, h) for _ in range(ndyes)]
        of.u[:], of.v[:], of.u_prev[:], of.v_prev[:] = init[:4]
        for i, od in enumerate(ods):
            od.dens[:], od.dens_prev[:] = init[4 + 2 * i], init[5 + 2 * i]
        for dt in dts:
            of.step(dt, visc)
            for od in ods:
                od.step((of.u, of.v), dt, diff)
        _oracle_cache[key] = [of.u, of.v, of.u_prev, of.v_prev] + [a for od in ods for a in (od.dens, od.dens_prev)]
    want = _oracle_cache[key]
    sim = fb.FluidSim(w, h)
    dyes = [fb.DensitySim(w, h) for _ in range(ndyes)]
    for a, f in zip(init[:4], range(4)):
        _lib.check(L.fruid_fluid_upload(sim._h, f, _lib.ptr(a)))
    for i, d in enumerate(dyes):
        _lib.check(L.fruid_density_upload(d._h, 0, _lib.ptr(init[4 + 2 * i])))
        _lib.check(L.fruid_density_upload(d._h, 1, _lib.ptr(init[5 + 2 * i])))
    for dt in dts:
        sim.step(dt, visc)
        fb.DensitySim.step_all(dyes, sim.uv(), dt, diff)
    got = [sim.uv()[0].numpy(), sim.uv()[1].numpy(), sim.uv_mut()[0].numpy(), sim.uv_mut()[1].numpy()]
    for d in dyes:
        got += [d.density().numpy(), d.density_mut().numpy()]
    names = ["u", "v", "u_prev", "v_prev"] + [f"{n}{i}" for i in range(ndyes) for n in ("dens", "dens_prev")]
    out = {}
    for n, g, wnt in zip(names, got, want):
        k = bits_bad(g, wnt, n)
        if k:
            out[n] = k
    return out


CASES = [
    ("golden37x29visc", lambda: case_golden("steps_37x29_visc.npz")),
    ("golden30x41", lambda: case_golden("steps_30x41_inviscid.npz")),
    ("rand96x80_visc_1dye_5steps", lambda: case_random(96, 80, 5, 1e-4, 1e-4, 1, 11)),
    ("rand250_4dyes_6steps_dtchange", lambda: case_random(250, 250, 6, 0.0, 0.0, 4, 12, True)),
    ("rand700x300_visc_2dyes_4steps", lambda: case_random(700, 300, 4, 2e-5, 1e-5, 2, 13)),
]

results = []
fails_safe = 0
combos = list(itertools.product((1, 0), (1, 0), (0, 1), (1, 0), (0, 1)))  # graphs, pdl, overlap, fork, unchained
if args.quick:
    combos = [c for c in combos if c[3] == 1 or c[2] == 0]
_lib.check(L.fruid_debug_set_tail_delay(0) if not debug_build else 0)
for graphs, pdl, overlap, fork, unchained in combos:
    if unchained and not pdl:
        continue  # the attribution only matters with programmatic launch on
    L.fruid_set_graphs_enabled(graphs)
    L.fruid_set_launch_overlap(pdl)
    L.fruid_debug_set_schedule(overlap | (fork << 1) | (unchained << 2))
    if debug_build:
        _lib.check(L.fruid_debug_set_tail_delay(args.delay_ns))
    for name, fn in CASES:
        if args.only and args.only not in name:
            continue
        for rep in range(args.repeat):
            bad = fn()
            print("run", graphs, pdl, overlap, fork, unchained, name, "bad" if bad else "ok", flush=True)
            rec = {"graphs": graphs, "pdl": pdl, "dens_overlap": overlap, "vel_fork": fork, "unchained": unchained,
                   "case": name, "rep": rep, "mismatching_cells": bad}
            results.append(rec)
            if bad:
                print("MISMATCH", json.dumps(rec), flush=True)
                while DETAILS:
                    print("   ", json.dumps(DETAILS.pop(0)), flush=True)
                if name.startswith("golden37"):
                    import ctypes
                    for cc in (np.float32(1.0) + np.float32(4.0) * np.float32(np.float32(np.float32(np.float32(0.05) * np.float32(v)) * np.float32(35)) * np.float32(27)) for v in (2e-4, 1e-4)):
                        o, z = (ctypes.c_int * 2)(), (ctypes.c_float * 2)()
                        L.fruid_debug_recip_mode(float(cc), ctypes.cast(o, ctypes.c_void_p), ctypes.cast(z, ctypes.c_void_p))
                        print("    recip", float(cc), list(o), [float(z[0]).hex(), float(z[1]).hex()], flush=True)
                if not unchained:
                    fails_safe += 1
L.fruid_debug_set_schedule(2)
L.fruid_set_graphs_enabled(1)
L.fruid_set_launch_overlap(1)
if debug_build:
    L.fruid_debug_set_tail_delay(0)
n_unchained_bad = sum(1 for r in results if r["unchained"] and r["mismatching_cells"])
summary = {"build": info, "delay_ns": args.delay_ns if debug_build else 0, "runs": len(results),
           "chained_mismatches": fails_safe, "unchained_mismatches": n_unchained_bad,
           "unchained_failing_schedules": sorted({(r["graphs"], r["dens_overlap"], r["vel_fork"], r["case"])
                                                  for r in results if r["unchained"] and r["mismatching_cells"]})}
print("ORDER PROBE", json.dumps(summary))
if args.json:
    with open(args.json, "w") as fh:
        json.dump({"summary": summary, "results": results}, fh, indent=1)
if args.assert_safe and fails_safe:
    sys.exit(1)
