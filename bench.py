#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (cell-steps/sec of the stable-fluids scene step) on B200.

A "step" is one SCENE STEP = 1x FluidSim::step (src/lib.rs:267) + D x DensitySim::step (src/lib.rs:226)
over one w x h grid with K Jacobi sweeps per lin_solve (src/lib.rs:52: 20); cell-steps/s = w*h*steps / seconds
(border cells included, like the constructor arguments).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config NAME]

--config (default `headline`, the line the driver records):
    headline    BASELINE.json configs[2]: 4096 x 4096, K=20 (40 pressure sweeps per step), D=1, Taylor-Green at CFL~2.
                With --gpus N: weak scaling, 4096 x 4096*N, y-slabs over NVLink peer memory.
    cfl64       the same at CFL~64 (scattered gather)
    viscous     the same with visc = diff = 1e-6 (general divisor in the three diffusion solves)
    scene250 / scene128   configs[0]: the reference viewer's scene (src/main.rs:36-52, 91-118): 250 x 250 (128 x 128),
                four dyes, per frame one injected velocity cell + four cleared dye sources
    inject1024  configs[1]: 1024 x 1024, D=1, the same injection law, 1000 steps by default
    ksweep      configs[3]: 4096 x 4096, K in {20, 40, 100, 200, 500}, D in {0, 1}; one line with a `sweep` list
    strong16k / weak16k   configs[4]: 16384 x 16384 (strong) or 16384 x 2048*N (weak) over N GPUs

Prints ONE JSON line on rank 0.  value = device-timed throughput with all fields resident in HBM; e2e = the same
metric through the public API with HOST buffers inside the timed region; roofline = the dominant kernel against the
measured HBM peak (with its actual DRAM traffic and FP32-pipe fractions beside the effective one); cpu_baseline = the
CPU oracle (faithful loop order, 1 thread, like the single-threaded reference) on a bounded sample.
--impl reference times that CPU path alone (the Rust reference itself cannot be built in this image: no rustc/cargo,
un-vendored deps) on a bounded sample whose true grid is printed.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT, VISC, DIFF, ITERS = 1e-2, 0.0, 0.0, 20  # dt, visc, diff of the reference scene (src/main.rs:110-112)
SEED = 0x5EED0001
FIELDS6 = ("u", "v", "u_prev", "v_prev", "dens", "dens_prev")


# ------------------------------------------------------------------------------------------
# synthetic scene: Taylor-Green velocity at CFL ~ 2 cells + hash noise,
# Gaussian dye blob + noise.  Deterministic in the GLOBAL cell index, so any slab of it can be
# generated independently (multi-GPU ranks see identical data).
# ------------------------------------------------------------------------------------------
def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    return z ^ (z >> np.uint64(31))


def noise(field: int, w: int, y0: int, y1: int) -> np.ndarray:
    """r in [-1, 1), exactly representable: ((splitmix64(seed ^ field<<56 ^ idx) >> 40) * 2^-24)*2 - 1."""
    with np.errstate(over="ignore"):
        idx = (np.arange(y0, y1, dtype=np.uint64)[:, None] * np.uint64(w) + np.arange(w, dtype=np.uint64)[None, :])
        z = _splitmix64(idx ^ np.uint64(SEED) ^ (np.uint64(field) << np.uint64(56)))
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)) * np.float32(2.0) - np.float32(1.0)


def scene_field(name: str, w: int, h: int, y0: int = 0, y1: int | None = None, cfl: float = 2.0) -> np.ndarray:
    """Rows [y0, y1) of one field of the global w x h scene, float32 (rows, w)."""
    y1 = h if y1 is None else y1
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    two_pi = np.float32(2 * np.pi)
    if name in ("u", "v"):
        amp = np.float32(cfl / (DT * (w - 2)))
        xh = (np.arange(w, dtype=np.float32) / np.float32(w))[None, :]
        yh = (np.arange(y0, y1, dtype=np.float32) / np.float32(h))[:, None]
        if name == "u":
            return f32(amp * np.sin(two_pi * xh) * np.cos(two_pi * yh) + np.float32(1e-3) * noise(0, w, y0, y1))
        return f32(-amp * np.cos(two_pi * xh) * np.sin(two_pi * yh) + np.float32(1e-3) * noise(1, w, y0, y1))
    if name == "dens":
        sig = np.float32(w / 16.0)
        dx = np.arange(w, dtype=np.float32)[None, :] - np.float32(w / 2)
        dy = np.arange(y0, y1, dtype=np.float32)[:, None] - np.float32(h / 2)
        return f32(np.exp(-(dx * dx + dy * dy) / (np.float32(2.0) * sig * sig)).astype(np.float32)
                   + np.float32(1e-3) * noise(2, w, y0, y1))
    # u_prev / v_prev / dens_prev are the per-step SOURCE fields (uv_mut / density_mut)
    if name == "u_prev":
        return f32(np.float32(1e-2) * noise(3, w, y0, y1))
    if name == "v_prev":
        return f32(np.float32(1e-2) * noise(4, w, y0, y1))
    if name == "dens_prev":
        return f32(np.float32(1e-2) * (noise(5, w, y0, y1) + np.float32(1.0)))
    if name == "zero":
        return np.zeros((y1 - y0, w), np.float32)
    raise KeyError(name)


def make_scene(w: int, h: int, y0: int = 0, y1: int | None = None, cfl: float = 2.0):
    """Rows [y0, y1) of the global w x h scene: dict of float32 (rows, w) arrays."""
    return {k: scene_field(k, w, h, y0, y1, cfl) for k in FIELDS6 + ("zero",)}


def bytes_per_cell_step(k: int = ITERS, dyes: int = 1) -> int:
    """Algorithmic HBM bytes per cell per scene step: (160 + 48K) + D*(28 + 12K); 188 + 60K at D = 1."""
    return (160 + 48 * k) + dyes * (28 + 12 * k)


def injection(frame_count: int, w: int, h: int):
    """(x, y, u_val, v_val) of the reference's per-frame velocity injection (src/main.rs:92-102), evaluated in f32."""
    F = np.float32
    t = F(frame_count) / F(12.0)
    x = F(w // 2) * (np.cos(t / F(5.0), dtype=F) + F(1.0))
    return min(int(x), w - 1), h // 2, float(F(-4500.0) * np.cos(t * F(3.0), dtype=F)), float(F(-4500.0) * np.sin(t * F(3.0), dtype=F))


class Workload:
    def __init__(self, name, w, h, iters=ITERS, dyes=1, cfl=2.0, visc=VISC, diff=DIFF, frame_inputs=False, steps=None,
                 what=""):
        self.name, self.w, self.h, self.iters, self.dyes, self.cfl = name, w, h, iters, dyes, cfl
        self.visc, self.diff, self.frame_inputs, self.steps, self.what = visc, diff, frame_inputs, steps, what

    def config(self, n_gpus):
        w, h = self.w, self.h
        fits = (6 + 2 * self.dyes) * w * h * 4 / max(n_gpus, 1) > (126 << 20)
        return {"workload": f"{w}x{h} scene step: FluidSim::step + {self.dyes}x DensitySim::step, K={self.iters} sweeps per "
                            f"lin_solve ({2 * self.iters} pressure sweeps/step), dt={DT}, visc={self.visc}, diff={self.diff} "
                            f"({self.what})",
                "name": self.name, "grid": [w, h], "jacobi_iters": self.iters, "dyes": self.dyes, "dt": DT,
                "visc": self.visc, "diff": self.diff,
                "init": ("reference scene: dye blobs, one injected velocity cell and cleared dye sources every frame "
                         "(src/main.rs:36-52,91-118)") if self.frame_inputs else
                        f"Taylor-Green velocity at CFL~{self.cfl:g} cells + hash noise; Gaussian dye; noisy source fields",
                "l2": (f"inputs larger than L2 ({6 + 2 * self.dyes} fields x {w * h * 4 / n_gpus / 2**20:.0f} MiB per GPU)" if fits else
                       f"L2-RESIDENT working set ({(6 + 2 * self.dyes) * w * h * 4 / 2**20:.1f} MiB): a latency / launch-bound "
                       f"configuration of the reference, not an HBM roofline case; no L2 flush between steps because a "
                       f"step consumes its predecessor's output"),
                "parallelism": f"slab{n_gpus}"}


def workload_from_args(args, n_gpus):
    c = args.config
    if c in ("headline", "cfl64", "viscous"):
        w = args.size or 4096
        h = args.rows or (w * n_gpus if n_gpus > 1 else w)
        return Workload(c, w, h, args.iters or ITERS, 1 if args.dyes is None else args.dyes,
                        cfl=64.0 if c == "cfl64" else 2.0, visc=1e-6 if c == "viscous" else VISC,
                        diff=1e-6 if c == "viscous" else DIFF,
                        what="BASELINE.json configs[2]" + ("" if c == "headline" else f", variant {c}"))
    if c in ("scene250", "scene128"):
        n = 250 if c == "scene250" else 128
        return Workload(c, n, n, args.iters or ITERS, 4 if args.dyes is None else args.dyes, frame_inputs=True, steps=100,
                        what="BASELINE.json configs[0], the reference viewer's scene headless")
    if c == "inject1024":
        return Workload(c, 1024, 1024, args.iters or ITERS, 1 if args.dyes is None else args.dyes, frame_inputs=True,
                        steps=1000, what="BASELINE.json configs[1]")
    if c == "ksweep":
        w = args.size or 4096
        return Workload(c, w, w, ITERS, 1, what="BASELINE.json configs[3]")
    if c == "strong16k":
        return Workload(c, 16384, args.rows or 16384, args.iters or ITERS, 1, what="BASELINE.json configs[4], strong scaling")
    if c == "weak16k":
        return Workload(c, 16384, args.rows or 2048 * n_gpus, args.iters or ITERS, 1,
                        what="BASELINE.json configs[4], weak scaling: 2048 rows per GPU")
    raise SystemExit(f"unknown --config {c}")


# ------------------------------------------------------------------------------------------
# clocks (sampled DURING the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int, enabled: bool = True, period_s: float = 0.05):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.period_s = period_s
        self._stop = threading.Event()
        self._thr = None
        try:
            if not enabled:
                raise RuntimeError("disabled")
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:  # NVML unavailable: report nulls rather than fail the bench
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period_s)

    def __enter__(self):
        if self._nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        return {"sm_mhz": (float(np.median(self.samples)) if self.samples else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU path (the oracle = restatement of the reference; used ONLY as baseline / reference arm)
# ------------------------------------------------------------------------------------------
def cpu_scene_steps(wl: Workload, rows: int, steps: int, warmup: int = 0, faithful: bool = True):
    """Times `steps` scene steps of the C oracle on a w x rows grid (a bounded sample of the w x h workload):
    faithful = the reference's loop order on ONE pinned thread (src/lib.rs is single-threaded); otherwise the
    cache-friendly order on every core (OpenMP; same bits, NOT how the reference runs).
    Returns (cell_steps_per_s, seconds_per_step, threads)."""
    import oracle
    oracle.build()
    w = wl.w
    threads = 1
    saved_affinity = None
    if faithful:
        try:
            saved_affinity = os.sched_getaffinity(0)
            os.sched_setaffinity(0, {sorted(saved_affinity)[0]})  # pin: the reference is single-threaded
        except Exception:
            saved_affinity = None
    else:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    order = oracle.FAITHFUL if faithful else oracle.FAST
    f = oracle.Fluid(w, rows, iters=wl.iters, order=order, openmp=not faithful)
    ds = [oracle.Density(w, rows, iters=wl.iters, order=order, openmp=not faithful) for _ in range(wl.dyes)]
    if not wl.frame_inputs:
        f.u[:], f.v[:] = scene_field("u", w, rows, cfl=wl.cfl), scene_field("v", w, rows, cfl=wl.cfl)
        f.u_prev[:], f.v_prev[:] = scene_field("u_prev", w, rows), scene_field("v_prev", w, rows)
        for d in ds:
            d.dens[:], d.dens_prev[:] = scene_field("dens", w, rows), scene_field("dens_prev", w, rows)
    else:
        for i, d in enumerate(ds):
            d.dens_prev[rows // 2, (i + 1) * w // 5 % w] = np.float32(80.0) * np.float32(w * rows)

    def step(i):
        if wl.frame_inputs:
            x, y, uv, vv = injection(i + 1, w, rows)
            f.u_prev[y, x], f.v_prev[y, x] = uv, vv
            if i > 0:
                for d in ds:
                    d.dens_prev.fill(0.0)
        f.step(DT, wl.visc)
        for d in ds:
            d.step((f.u, f.v), DT, wl.diff)
    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    sec = (time.perf_counter() - t0) / max(steps, 1)
    if saved_affinity is not None:
        try:
            os.sched_setaffinity(0, saved_affinity)
        except Exception:
            pass
    return w * rows / sec, sec, threads


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; 1 thread because src/lib.rs is
    single-threaded), same metric, on a bounded sample: the first `rows` rows' worth of the workload, with `rows`
    sized from a calibration step so that warm-up + steps take about --ref-seconds.  The grid actually run is what
    config.grid and ms_per_step state."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload_from_args(args, 1)
    if wl.name == "ksweep":
        wl.iters = ITERS
    steps = args.steps if args.steps is not None else (wl.steps or 20)
    total = steps + args.warmup
    cal_rows = min(wl.h, 128)
    _, cal_sec, _ = cpu_scene_steps(wl, cal_rows, 1)
    per_row = cal_sec / cal_rows
    rows = int(args.ref_seconds / max(total, 1) / per_row)
    rows = wl.h if rows >= wl.h else max(64, rows // 64 * 64)
    rows = min(rows, wl.h)
    value, sec, _ = cpu_scene_steps(wl, rows, steps, args.warmup)
    omp = None
    if not args.no_omp:
        ov, osec, othreads = cpu_scene_steps(wl, rows, 1, 1, faithful=False)
        omp = {"value": ov, "unit": "cell-steps/s", "cores": othreads, "ms_per_step": osec * 1e3,
               "note": "NOT how the reference runs: cache-friendly loop order + OpenMP on every core (same bits); context only"}
    frac = rows / wl.h
    sample = (f"each step = 1 scene step (K={wl.iters}, D={wl.dyes}) of the faithful-order C oracle on a {wl.w}x{rows} grid "
              f"({'the whole' if rows == wl.h else f'{frac:.1%} of the'} {wl.w}x{wl.h} workload), 1 thread pinned "
              f"(src/lib.rs is single-threaded); the Rust reference cannot be built here")
    cfg = wl.config(1)
    cfg["grid"] = [wl.w, rows]
    cfg["sampled_from_grid"] = [wl.w, wl.h]
    if rows != wl.h:
        cfg["workload"] = f"the first {rows} rows ({wl.w}x{rows} grid) of: " + cfg["workload"]
    line = {
        "impl": "reference", "metric": "cell-steps/sec", "value": value, "unit": "cell-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "cell-steps/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cores_available": os.cpu_count(), "all_cores_variant": omp,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# B200 path
# ------------------------------------------------------------------------------------------
class Rig:
    """One velocity handle + D dye handles on one GPU, driven through the C ABI with device-resident fields."""

    def __init__(self, fb, wl: Workload, fused=None):
        from fruid_b200 import _lib
        self._lib, self.L, self.wl = _lib, _lib.lib, wl
        w, h = wl.w, wl.h
        self.sim = fb.FluidSim(w, h)
        self.dyes = [fb.DensitySim(w, h) for _ in range(wl.dyes)]
        for o in [self.sim] + self.dyes:
            o.set_iterations(wl.iters)
            if fused is not None:
                o.set_fused_sweeps(fused)
        self.handles = (_lib._vp * max(len(self.dyes), 1))(*[d._h for d in self.dyes])
        self.frame = 0
        self.load()

    def load(self):
        _lib, L, wl = self._lib, self.L, self.wl
        w, h = wl.w, wl.h
        if not wl.frame_inputs:
            for f, name in enumerate(("u", "v", "u_prev", "v_prev")):
                _lib.check(L.fruid_fluid_upload(self.sim._h, f, scene_field(name, w, h, cfl=wl.cfl)))
            for d in self.dyes:
                _lib.check(L.fruid_density_upload(d._h, 0, scene_field("dens", w, h)))
                _lib.check(L.fruid_density_upload(d._h, 1, scene_field("dens_prev", w, h)))
        else:  # main.rs:42-52: dye blobs, one initial step at dt = 0.1
            inten = float(np.float32(80.0) * np.float32(w * h))
            place = [(w // 5, inten), (3 * w // 5, inten), (4 * w // 5, inten), (2 * w // 5, inten / 100.0)]
            for d, (x, val) in zip(self.dyes, place):
                _lib.check(L.fruid_density_inject(d._h, 1, x, h // 2, val))
            _lib.check(L.fruid_fluid_step(self.sim._h, 0.1, 0.0))
            self._dyes(0.1, 0.0)

    def _dyes(self, dt, diff):
        _lib, L = self._lib, self.L
        if len(self.dyes) == 1:
            _lib.check(L.fruid_density_step_dev(self.dyes[0]._h, self.sim._h, dt, diff))
        elif self.dyes:
            _lib.check(L.fruid_density_step_dev_batch(self.handles, len(self.dyes), self.sim._h, dt, diff))

    def step(self, visc=None, diff=None):
        _lib, L, wl = self._lib, self.L, self.wl
        if wl.frame_inputs:  # main.rs:91-108
            self.frame += 1
            x, y, uv, vv = injection(self.frame, wl.w, wl.h)
            _lib.check(L.fruid_fluid_inject(self.sim._h, 2, x, y, uv))
            _lib.check(L.fruid_fluid_inject(self.sim._h, 3, x, y, vv))
            for d in self.dyes:
                _lib.check(L.fruid_density_fill(d._h, 1, 0.0))
        _lib.check(L.fruid_fluid_step(self.sim._h, DT, wl.visc if visc is None else visc))
        self._dyes(DT, wl.diff if diff is None else diff)

    def set_iterations(self, k):
        for o in [self.sim] + self.dyes:
            o.set_iterations(k)

    def stream(self, torch):
        return torch.cuda.ExternalStream(int(self.L.fruid_fluid_stream(self.sim._h)))


def timed(torch, rig, stream, steps, warmup, sampler=None):
    for _ in range(warmup):
        rig.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        rig.step()
    e1.record(stream)   # dye work is ordered before this by the cross-stream events
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def roofline_of(stats, wl, scene_steps, value, clocks, in_step_us_per_sweep=None):
    """The dominant kernel (the fused Jacobi kernel).  Its launch duration comes from CUDA events over timed regions of
    the real step loop (graph replay, programmatic launch, the density stream running beside the velocity stream): the
    same loop is timed at K and at 2K sweeps per lin_solve and the difference is divided by the extra launches
    (`in_step_us_per_sweep`).  The per-kernel pass that brackets every launch with its own events (profiling mode: direct
    launches, nothing overlapped) gives the kernel's share of the step and the other kernels' times; its launch time is
    kept as `avg_launch_us_isolated`."""
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    w, h = wl.w, wl.h
    tot_ms = sum(t for _, t in stats.values()) or 1.0
    jac = {k: v for k, v in stats.items() if k.startswith("k_jacobi")}
    dom_name = max(jac, key=lambda k: jac[k][1]) if jac else max(stats, key=lambda k: stats[k][1])
    dom_launches, dom_ms = stats[dom_name]
    sweeps_per_launch = (4 + wl.dyes) * wl.iters * scene_steps / dom_launches
    alg_bytes = 12.0 * w * h * sweeps_per_launch      # 12 B/cell/sweep
    iso_s = dom_ms * 1e-3 / dom_launches
    avg_s = in_step_us_per_sweep * 1e-6 * sweeps_per_launch if in_step_us_per_sweep else iso_s
    achieved = alg_bytes / avg_s / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom_name)
        if traffic is not None and (w, h) != (4096, 4096):
            traffic = traffic * (w * h) / (4096.0 * 4096.0)  # the ncu capture was taken at 4096 x 4096
    except Exception:
        pass
    mhz = (clocks or {}).get("sm_mhz") or 1900.0
    # 5 FP32 lane-operations per cell and sweep (three adds of the stencil sum, x0 + a*sum, the scaling); 148 SMs x 128 lanes
    fp32 = 5.0 * w * h * sweeps_per_launch / (148 * 128 * mhz * 1e6 * avg_s)
    return {"bound": "fp32-issue (the fused kernel is not HBM-bound: see dram_frac_actual)", "kernel": dom_name,
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch"
                              + (", scaled by cells" if (w, h) != (4096, 4096) else "") + ")",
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
            "frac_note": "EFFECTIVE bandwidth: algorithmic (unfused) bytes of the sweeps one launch performs / its CUDA-event "
                         "time; > 1 is possible because the sweeps are fused in registers",
            "fused_floor_frac": 12.0 * w * h / avg_s / 1e9 / peak,
            "fused_floor_note": "12 B/cell once per launch (read x, read x0, write x') / time / peak: how far the launch is "
                                "from the HBM floor of a perfectly fused kernel",
            "dram_frac_actual": (traffic / avg_s / 1e9 / peak) if traffic else None,
            "fp32_pipe_frac": fp32,
            "fp32_pipe_note": f"5 FP32 lane-ops per cell-sweep / (148 SMs x 128 lanes x {mhz:.0f} MHz), halo recompute excluded",
            "sweeps_per_launch": sweeps_per_launch, "avg_launch_us": avg_s * 1e6, "avg_launch_us_isolated": iso_s * 1e6,
            "launch_time_method": ("CUDA events over the timed step loop at K and 2K sweeps per lin_solve: (t(2K) - t(K)) / extra "
                                   "launches" if in_step_us_per_sweep else "CUDA events around every launch (profiling pass)"),
            "algorithmic_bytes_per_launch": alg_bytes, "kernel_share_of_step": dom_ms / tot_ms,
            "kernels_ms_per_step": {k: round(t / scene_steps, 4) for k, (n, t) in sorted(stats.items())},
            "kernels_launches_per_step": {k: n / scene_steps for k, (n, t) in sorted(stats.items())},
            "step_effective_gbs": bytes_per_cell_step(wl.iters, wl.dyes) * value / 1e9,
            "step_effective_frac": bytes_per_cell_step(wl.iters, wl.dyes) * value / 1e9 / peak}


def run_b200(args):
    import torch
    import __graft_entry__
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; fruid_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if rank == 0:
        __graft_entry__.build()
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    import fruid_b200 as fb
    from fruid_b200 import _lib

    wl = workload_from_args(args, world)
    if world > 1:
        return run_b200_multi(args, wl, fb, dist, rank, world, local)
    steps = args.steps if args.steps is not None else (wl.steps or 100)
    if wl.name == "ksweep":
        return run_ksweep(args, wl, fb, torch, local)

    rig = Rig(fb, wl, args.fused)
    stream = rig.stream(torch)

    # ---- device-resident throughput ("value") ----
    for _ in range(args.warmup):
        rig.step()
    torch.cuda.synchronize()
    n0 = fb.launch_count()
    with ClockSampler(local, period_s=0.01) as clk:
        ms = timed(torch, rig, stream, steps, 0)
    launches = fb.launch_count() - n0
    ms_per_step = ms / steps
    value = wl.w * wl.h * steps / (ms * 1e-3)
    clocks = clk.summary()

    # ---- sustained figure: the same loop for about --sustain-seconds ----
    sustained = None
    if args.sustain_seconds > 0 and not wl.frame_inputs:
        n_s = max(steps, int(args.sustain_seconds * 1e3 / ms_per_step))
        with ClockSampler(local, period_s=0.05) as clk_s:
            ms_s = timed(torch, rig, stream, n_s, 0)
        sustained = {"value": wl.w * wl.h * n_s / (ms_s * 1e-3), "unit": "cell-steps/s", "steps": n_s, "seconds": ms_s * 1e-3,
                     "ms_per_step": ms_s / n_s, "clocks": clk_s.summary()}

    # ---- the fused Jacobi kernel's time per sweep inside the real step loop: the same loop at 2K sweeps ----
    in_step = None
    if not wl.frame_inputs:
        n2 = max(3, min(steps, 10))
        ms_k = timed(torch, rig, stream, n2, 1) / n2
        rig.set_iterations(2 * wl.iters)
        ms_2k = timed(torch, rig, stream, n2, 2) / n2
        rig.set_iterations(wl.iters)
        timed(torch, rig, stream, 1, 1)
        in_step = (ms_2k - ms_k) * 1e3 / ((4 + wl.dyes) * wl.iters)   # us per sweep of one field

    # ---- per-kernel timing pass (shares of the step, the other kernels) ----
    nprof = max(2, min(steps, 5))
    _lib.profile_enable(True)
    for _ in range(nprof):
        rig.step()
    stats = _lib.profile_report()
    _lib.profile_enable(False)
    roofline = roofline_of(stats, wl, nprof, value, clocks, in_step)

    # ---- variants on the headline workload ----
    variants = {}
    if wl.name == "headline" and not args.no_variants:
        def timed_variant(visc, diff, n):
            for _ in range(3):
                rig.step(visc, diff)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(n):
                rig.step(visc, diff)
            a1.record(stream)
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / n
        vms = timed_variant(1e-6, 1e-6, max(2, min(steps, 10)))
        variants["viscous"] = {"visc": 1e-6, "diff": 1e-6, "ms_per_step": vms, "value": wl.w * wl.h / (vms * 1e-3),
                               "unit": "cell-steps/s",
                               "note": "same scene with visc = diff = 1e-6 (SURVEY.md 8d config 1 variant B): the diffusion "
                                       "solves divide by c = 1 + 4a through the verified reciprocal path"}
        del rig
        wl64 = Workload("cfl64", wl.w, wl.h, wl.iters, wl.dyes, cfl=64.0)
        rig64 = Rig(fb, wl64, args.fused)
        ms64 = timed(torch, rig64, rig64.stream(torch), max(2, min(steps, 10)), 3) / max(2, min(steps, 10))
        _lib.profile_enable(True)
        for _ in range(2):
            rig64.step()
        st64 = _lib.profile_report()
        _lib.profile_enable(False)
        variants["cfl64"] = {"cfl": 64.0, "ms_per_step": ms64, "value": wl.w * wl.h / (ms64 * 1e-3), "unit": "cell-steps/s",
                             "advect_us": {k: round(t / n * 1e3, 2) for k, (n, t) in st64.items() if k.startswith("k_advect")},
                             "note": "same scene at CFL~64 cells (scattered gather; SURVEY.md 8d config 2, second run)"}
        del rig64
    else:
        del rig

    # ---- end to end through the public API with HOST buffers ----
    e2e = run_e2e_frames(args, fb, torch, wl) if wl.frame_inputs else run_e2e(args, fb, torch, wl)

    # ---- CPU baseline beside it (bounded sample) ----
    cpu = None
    if not args.no_cpu:
        rows = min(1024, wl.h)
        cv, csec, _ = cpu_scene_steps(wl, rows, 1)
        cpu = {"value": cv, "unit": "cell-steps/s", "cores": 1, "kind": "port",
               "sample": f"1 scene step (K={wl.iters}, D={wl.dyes}) of the faithful-order C oracle on a {wl.w}x{rows} grid "
                         f"({csec:.1f} s), 1 thread pinned (the reference is single-threaded); "
                         f"host has {os.cpu_count()} cores"}

    line = {"metric": "cell-steps/sec", "value": value, "unit": "cell-steps/s", "n_gpus": 1, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wl.config(1),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "impl": "b200", "sustained": sustained}
    if variants:
        line["variant_viscous"] = variants.get("viscous")
        line["variant_cfl64"] = variants.get("cfl64")
    print(json.dumps(line), flush=True)


def run_ksweep(args, wl, fb, torch, local):
    """BASELINE.json configs[3]: K in {20, 40, 100, 200, 500} at 4096^2, D in {0, 1}: per-sweep time of the fused kernel,
    effective (algorithmic) and actual GB/s per K and fused depth."""
    from fruid_b200 import _lib
    steps = args.steps if args.steps is not None else 5
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    sweep, first = [], None
    n0 = fb.launch_count()
    for dyes in (1, 0):
        w1 = Workload("ksweep", wl.w, wl.h, ITERS, dyes, what="BASELINE.json configs[3]")
        rig = Rig(fb, w1, args.fused)
        stream = rig.stream(torch)
        for k in (20, 40, 100, 200, 500):
            rig.set_iterations(k)
            w1.iters = k
            with ClockSampler(local, period_s=0.01) as clk:
                ms = timed(torch, rig, stream, steps, 3)
            value = wl.w * wl.h * steps / (ms * 1e-3)
            _lib.profile_enable(True)
            for _ in range(2):
                rig.step()
            stats = _lib.profile_report()
            _lib.profile_enable(False)
            r = roofline_of(stats, w1, 2, value, clk.summary())
            rec = {"K": k, "dyes": dyes, "ms_per_step": ms / steps, "value": value,
                   "us_per_sweep": r["avg_launch_us"] / r["sweeps_per_launch"], "sweeps_per_launch": r["sweeps_per_launch"],
                   "jacobi_effective_gbs": r["achieved"], "jacobi_effective_frac": r["frac"],
                   "jacobi_dram_frac_actual": r["dram_frac_actual"], "fp32_pipe_frac": r["fp32_pipe_frac"],
                   "step_effective_gbs": r["step_effective_gbs"], "step_effective_frac": r["step_effective_frac"],
                   "jacobi_share_of_step": r["kernel_share_of_step"], "clocks": clk.summary()}
            sweep.append(rec)
            if first is None:
                first = (rec, r, clk.summary())
        del rig
    launches = fb.launch_count() - n0
    rec, r, clocks = first
    cfg = wl.config(1)
    cfg["K_values"] = [20, 40, 100, 200, 500]
    line = {"metric": "cell-steps/sec", "value": rec["value"], "unit": "cell-steps/s", "n_gpus": 1, "steps": steps,
            "warmup": 3, "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": None,
            "gpu_launches": int(launches), "roofline": r, "cpu_baseline": None, "impl": "b200", "sweep": sweep,
            "note": "value / ms_per_step / roofline are the K=20, D=1 entry; `sweep` holds every (K, D); peak " + str(peak)}
    print(json.dumps(line), flush=True)


def run_e2e(args, fb, torch, wl):
    """Scene step through the C ABI with HOST buffers: every step the three source fields (u_prev, v_prev,
    dens_prev) come from page-locked host arrays (h2d) and the resulting density goes back to a host array
    (d2h).  The transfers use the ABI's asynchronous variants (fruid_*_upload_async / _download_async): the
    sources of step i+1 travel while step i computes and the density of step i travels while step i+1
    computes.  Wall clock over n steps including every copy, synchronised on both sides; the serial
    (blocking upload / step / download) figure is reported beside it."""
    from fruid_b200 import _lib
    L = _lib.lib
    w, h = wl.w, wl.h
    sim, den = fb.FluidSim(w, h), fb.DensitySim(w, h)
    for o in (sim, den):
        o.set_iterations(wl.iters)
        if args.fused is not None:
            o.set_fused_sweeps(args.fused)
    for f, name in enumerate(("u", "v")):
        _lib.check(L.fruid_fluid_upload(sim._h, f, scene_field(name, w, h, cfl=wl.cfl)))
    _lib.check(L.fruid_density_upload(den._h, 0, scene_field("dens", w, h)))
    # two sets of host source arrays and two result arrays (what a double-buffering caller keeps)
    host = [{k: scene_field(k, w, h) for k in ("u_prev", "v_prev", "dens_prev")} for _ in range(2)]
    out = [np.empty((h, w), np.float32) for _ in range(2)]
    arrays = [a for hs in host for a in hs.values()] + out
    pinned = [a for a in arrays if L.fruid_host_register(a.ctypes.data, a.nbytes) == 0]
    nbytes = w * h * 4

    def push(i):
        hs = host[i & 1]
        _lib.check(L.fruid_fluid_upload_async(sim._h, _lib.U_PREV, _lib.ptr(hs["u_prev"])))
        _lib.check(L.fruid_fluid_upload_async(sim._h, _lib.V_PREV, _lib.ptr(hs["v_prev"])))
        _lib.check(L.fruid_density_upload_async(den._h, _lib.DENS_PREV, _lib.ptr(hs["dens_prev"])))

    def pipelined(n):
        push(0)
        for i in range(n):
            _lib.check(L.fruid_fluid_step(sim._h, DT, wl.visc))              # consumes the staged sources
            _lib.check(L.fruid_density_step_dev(den._h, sim._h, DT, wl.diff))
            if i + 1 < n:
                push(i + 1)                                                # h2d behind step i's kernels
            _lib.check(L.fruid_density_download_async(den._h, _lib.DENS, _lib.ptr(out[i & 1])))  # d2h behind step i+1
        _lib.check(L.fruid_fluid_sync(sim._h))
        _lib.check(L.fruid_density_sync(den._h))

    def serial(n):
        for i in range(n):
            hs = host[i & 1]
            _lib.check(L.fruid_fluid_upload(sim._h, _lib.U_PREV, _lib.ptr(hs["u_prev"])))
            _lib.check(L.fruid_fluid_upload(sim._h, _lib.V_PREV, _lib.ptr(hs["v_prev"])))
            _lib.check(L.fruid_density_upload(den._h, _lib.DENS_PREV, _lib.ptr(hs["dens_prev"])))
            _lib.check(L.fruid_fluid_step(sim._h, DT, wl.visc))
            _lib.check(L.fruid_density_step_dev(den._h, sim._h, DT, wl.diff))
            _lib.check(L.fruid_density_download(den._h, _lib.DENS, _lib.ptr(out[i & 1])))

    n = max(2, min(args.steps if args.steps is not None else 10, 10))
    res = {}
    for name, fn in (("pipelined", pipelined), ("serial", serial)):
        fn(2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn(n)
        torch.cuda.synchronize()
        res[name] = w * h * n / (time.perf_counter() - t0)
    ok = bool(np.isfinite(out[(n - 1) & 1][:: max(1, h // 64), :: max(1, w // 64)]).all())
    for a in pinned:
        L.fruid_host_unregister(a.ctypes.data)
    return {"value": res["pipelined"], "unit": "cell-steps/s", "h2d_bytes_per_step": 3 * nbytes,
            "d2h_bytes_per_step": nbytes, "steps": n, "finite": ok, "pinned_host": len(pinned) == len(arrays),
            "serial_value": res["serial"],
            "note": "per step: h2d of the u_prev, v_prev and dens_prev source fields, FluidSim::step + DensitySim::step, "
                    "d2h of density(); C ABI with page-locked host buffers; value = asynchronous (overlapped) transfers, "
                    "serial_value = blocking upload/step/download"}


def run_e2e_frames(args, fb, torch, wl):
    """The reference's frame through the compiled crate-API mirror (include/fruid.hpp, the C++ twin of rust/src/lib.rs),
    literally main.rs:94-124: one-cell writes through uv_mut(), density_mut().data_mut().fill(0) on the four dyes,
    sim.step, the dyes' step(sim.uv(), ..), then -- what draw_density (main.rs:146-183) needs -- the RGB of every cell
    computed on the device and copied to the host.  Run by fruid_b200/lib/frame_e2e (fruid_b200/host/frame_e2e.cpp);
    the host <-> device bytes are what the mirror actually moved per frame (counted where it calls the C ABI).  The same
    frame through the Python mirror (fruid_b200.api) is reported beside it."""
    import subprocess
    from fruid_b200 import _lib
    from fruid_b200.build import FRAME_E2E
    from fruid_b200.scene import Scene
    L = _lib.lib
    w, h = wl.w, wl.h
    n = args.steps if args.steps is not None else (wl.steps or 100)
    n = max(2, min(n, 1000))
    out = {"unit": "cell-steps/s", "steps": n}
    if wl.dyes == 4:
        r = subprocess.run([FRAME_E2E, str(w), str(h), str(n), "5", "1"], capture_output=True, text=True, timeout=600)
        cpp = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stdout + r.stderr}
        r2 = subprocess.run([FRAME_E2E, str(w), str(h), str(n), "5", "0"], capture_output=True, text=True, timeout=600)
        cpp_unbatched = json.loads(r2.stdout.strip().splitlines()[-1]) if r2.returncode == 0 else {"error": r2.stdout + r2.stderr}
        if "us_per_frame" in cpp:
            out.update({"value": w * h / (cpp["us_per_frame"] * 1e-6), "us_per_frame": cpp["us_per_frame"],
                        "h2d_bytes_per_step": cpp["h2d_bytes_per_frame"], "d2h_bytes_per_step": cpp["d2h_bytes_per_frame"],
                        "finite": cpp["finite"], "cpp_mirror": cpp, "cpp_mirror_one_call_per_dye": cpp_unbatched})
        else:
            out["cpp_mirror"] = cpp
    # the Python mirror (what the tests drive)
    s = Scene(fb.FluidSim, fb.DensitySim, w, h, n_dyes=wl.dyes)
    rgb = np.empty((h, w, 3), np.float32)
    four = len(s.dyes) == 4

    def frame():
        s.frame(DT, wl.visc, wl.diff)
        if four:
            c, m, y, k = (d._h for d in s.dyes)
            _lib.check(L.fruid_density_colors(c, m, y, k, rgb.reshape(h, 3 * w)))
        else:
            s.dyes[0].density().numpy()
    npy = max(2, min(n, 200))
    for _ in range(3):
        frame()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(npy):
        frame()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    out["python_mirror_us_per_frame"] = sec / npy * 1e6
    if "value" not in out:  # (not the four-dye scene: only the Python mirror ran)
        out.update({"value": w * h * npy / sec, "us_per_frame": sec / npy * 1e6, "h2d_bytes_per_step": 8,
                    "d2h_bytes_per_step": w * h * 4, "finite": True})
    out["note"] = ("crate-API mirror, the literal main.rs:94-124 frame: uv_mut()[pos] = .. (2 cells, forwarded as injections), "
                   "density_mut().data_mut().fill(0) on every dye (device-side fill, nothing uploaded), FluidSim::step, the dyes' "
                   "step(sim.uv(), ..) on the device-resident velocity (batched; `cpp_mirror_one_call_per_dye` = one call per "
                   "dye), then the frame's RGB (device draw_density colour map) copied to the host; wall clock per frame, "
                   "synchronous like the viewer's loop; value = the compiled C++ mirror, python_mirror_us_per_frame = the same "
                   "frame through fruid_b200.api")
    return out


def pin_to_gpu_numa(torch, local):
    """Run this rank's host thread -- and therefore first-touch its host buffers -- on the NUMA node its GPU hangs off
    (sysfs: the PCI device's numa_node and that node's cpulist, intersected with the CPUs this process may use).
    Returns what was found, for the e2e line; a container that exposes one node only has nothing to choose."""
    info = {"gpu_numa_node": None, "cpus_allowed": None, "cpus_on_gpu_node": None, "pinned": False}
    try:
        allowed = os.sched_getaffinity(0)
        info["cpus_allowed"] = len(allowed)
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info["gpu_numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        mine = cpus & allowed
        info["cpus_on_gpu_node"] = len(mine)
        if mine and mine != allowed:
            os.sched_setaffinity(0, mine)
            info["pinned"] = True
    except Exception as e:  # no sysfs, no such attribute: leave the scheduler alone
        info["error"] = f"{type(e).__name__}: {e}"[:120]
    return info


def run_b200_multi(args, wl, fb, dist, rank, world, local):
    """N GPUs of one node, one process per GPU: the scene is y-slab decomposed (rank r owns rows
    [r*h/N, (r+1)*h/N) plus 22 halo rows of each neighbour); halo rows and advection gathers move
    over NVLink peer memory inside the library's kernels.  WEAK scaling by default: w x (w*N) cells, i.e. the
    single-GPU workload per GPU.  Time = max over ranks of the CUDA-event time of the K steps,
    bracketed by barriers.  After the measurements the slab result is compared with a single-GPU run of the
    same global grid (`parity_vs_1gpu`)."""
    import torch
    from fruid_b200 import _lib
    from fruid_b200.dist import SlabDensitySim, SlabFluidSim
    L = _lib.lib
    w, h = wl.w, wl.h
    numa = pin_to_gpu_numa(torch, local)   # before any host buffer of this rank is touched
    steps = args.steps if args.steps is not None else 100
    sim = SlabFluidSim(w, h, rank, world)
    dens = [SlabDensitySim(w, h, rank, world) for _ in range(wl.dyes)]
    den = dens[0]
    for o in [sim] + dens:
        o.set_iterations(wl.iters)

    def load(sim, dens):
        for f, name in enumerate(("u", "v", "u_prev", "v_prev")):
            sim.upload(f, scene_field(name, w, h, sim.y0, sim.y1, cfl=wl.cfl))
        for d in dens:
            d.upload(0, scene_field("dens", w, h, sim.y0, sim.y1))
            d.upload(1, scene_field("dens_prev", w, h, sim.y0, sim.y1))
    load(sim, dens)
    stream = torch.cuda.ExternalStream(int(L.fruid_fluid_stream(sim._h)))

    def scene_step():
        sim.step(DT, wl.visc)
        for d in dens:
            d.step(sim, DT, wl.diff)

    def sync_all():
        sim.sync()
        for d in dens:
            d.sync()

    for _ in range(args.warmup):
        scene_step()
    sync_all()
    n0 = fb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # NVML is sampled on rank 0 only and slowly: concurrent NVML queries from 8 processes serialise
    # in the driver and stall kernel launches.  The sampler starts BEFORE the barrier so that every
    # rank enters the timed region together.
    with ClockSampler(local, enabled=(rank == 0), period_s=0.2) as clk:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t_host = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            scene_step()
        e1.record(stream)
        t_enq = time.perf_counter() - t_host
        sync_all()
        torch.cuda.synchronize()
    dist.barrier()
    launches = fb.launch_count() - n0
    mine_ms = e0.elapsed_time(e1)
    ms = torch.tensor([mine_ms], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    per_rank = [None] * world
    dist.all_gather_object(per_rank, (round(mine_ms / steps, 4), round(t_enq * 1e3 / steps, 4)))
    value = w * h * steps / (ms * 1e-3)

    # per-kernel timing pass (rank 0's view; includes the time spent waiting for peers)
    _lib.profile_enable(True)
    for _ in range(3):
        scene_step()
    stats = _lib.profile_report()
    _lib.profile_enable(False)
    dist.barrier()
    kernels_ms = {k: [n // 3, round(t / 3, 4)] for k, (n, t) in sorted(stats.items())}
    compute_ms = round(sum(t for k, (n, t) in stats.items() if not k.startswith("k_slab")) / 3, 4)
    comm_ms = round(sum(t for k, (n, t) in stats.items() if k.startswith("k_slab")) / 3, 4)
    per_rank_split = [None] * world
    dist.all_gather_object(per_rank_split, (compute_ms, comm_ms))

    # end to end: every step each rank uploads its rows of the three source fields and downloads its rows
    # of the density (host <-> device inside the timed region), wall clock, max over ranks.  Like the
    # single-GPU leg: asynchronous transfers that overlap the neighbouring steps (value) and the blocking
    # upload / step / download sequence (serial_value).
    rows = sim.y1 - sim.y0
    e2e = None
    if not args.no_e2e:
        host = [{k: scene_field(k, w, h, sim.y0, sim.y1) for k in ("u_prev", "v_prev", "dens_prev")} for _ in range(2)]
        out = [np.empty((rows, w), np.float32) for _ in range(2)]
        arrays = [a for hs in host for a in hs.values()] + out
        pinned = [a for a in arrays if L.fruid_host_register(a.ctypes.data, a.nbytes) == 0]

        def push(i):
            hs = host[i & 1]
            _lib.check(L.fruid_fluid_upload_async(sim._h, _lib.U_PREV, _lib.ptr(hs["u_prev"])))
            _lib.check(L.fruid_fluid_upload_async(sim._h, _lib.V_PREV, _lib.ptr(hs["v_prev"])))
            _lib.check(L.fruid_density_upload_async(den._h, _lib.DENS_PREV, _lib.ptr(hs["dens_prev"])))

        def pipelined(n):
            push(0)
            for i in range(n):
                scene_step()                                                   # consumes the staged sources
                if i + 1 < n:
                    push(i + 1)                                                # h2d behind step i's kernels
                _lib.check(L.fruid_density_download_async(den._h, _lib.DENS, _lib.ptr(out[i & 1])))  # d2h behind step i+1
            sync_all()

        def serial(n):
            for i in range(n):
                hs = host[i & 1]
                sim.upload(_lib.U_PREV, hs["u_prev"])
                sim.upload(_lib.V_PREV, hs["v_prev"])
                den.upload(_lib.DENS_PREV, hs["dens_prev"])
                scene_step()
                _lib.check(L.fruid_density_download(den._h, _lib.DENS, _lib.ptr(out[i & 1])))

        n = max(2, min(steps, 10))
        res, h2d_gbs = {}, None
        for name, fn in (("pipelined", pipelined), ("serial", serial)):
            fn(2)
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            fn(n)
            torch.cuda.synchronize(); dist.barrier()
            sec = torch.tensor([time.perf_counter() - t0], device="cuda")
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
            res[name] = w * h * n / float(sec.item())
        # this rank's raw host -> device rate with every rank copying at once (the PCIe / host-memory ceiling of e2e)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for i in range(4):
            sim.upload(_lib.U_PREV, host[i & 1]["u_prev"])
        my_h2d = 4 * rows * w * 4 / (time.perf_counter() - t0) / 1e9
        all_h2d = [None] * world
        dist.all_gather_object(all_h2d, round(my_h2d, 2))
        all_numa = [None] * world
        dist.all_gather_object(all_numa, numa)
        for a in pinned:
            L.fruid_host_unregister(a.ctypes.data)
        e2e = {"value": res["pipelined"], "unit": "cell-steps/s", "h2d_bytes_per_step": 3 * w * h * 4,
               "d2h_bytes_per_step": w * h * 4, "steps": n, "serial_value": res["serial"],
               "pinned_host": len(pinned) == len(arrays), "per_rank_h2d_gbs_all_ranks_copying": all_h2d,
               "pcie_ceiling_cell_steps_per_s": sum(all_h2d) * 1e9 / 16.0, "numa_per_rank": all_numa,
               "note": "whole-job bytes; each rank moves its own rows (h2d of 3 source fields, d2h of density) every step; "
                       "value = asynchronous (overlapped) transfers, serial_value = blocking upload/step/download; "
                       "pcie_ceiling = summed measured h2d rate / 16 B per cell-step (12 up + 4 down share the links): what "
                       "e2e could reach if the kernels cost nothing"}

    # ---- multi-GPU parity: a fresh slab run against a single-GPU run of the SAME global grid, SHA-256 per rank ----
    parity = "skipped"
    if not args.no_parity:
        del sim, dens, den
        parity = slab_parity(wl, fb, dist, rank, world, steps=3)

    if rank == 0:
        cfg = wl.config(world)
        cfg["workload"] = (f"{w}x{h} scene step y-slab decomposed over {world} GPUs ({w}x{h // world} cells per GPU"
                           + (f", weak scaling of the {w}x{w} single-GPU workload" if wl.name == "headline" and not args.rows else "")
                           + f"); FluidSim::step + {wl.dyes}x DensitySim::step, K={wl.iters} ({wl.what})")
        cfg["halo_rows"] = 22
        peak = float(load_peaks().get("hbm_gbs", 6650.0))
        bpc = bytes_per_cell_step(wl.iters, wl.dyes)
        scaling = "weak" if (wl.name in ("headline", "weak16k", "cfl64", "viscous") and not args.rows) else "strong"
        line = {"metric": "cell-steps/sec", "value": value, "unit": "cell-steps/s", "n_gpus": world, "steps": steps,
                "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg, "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches),
                "parity_vs_1gpu": parity,
                "roofline": {"bound": "fp32-issue", "kernel": "whole step", "achieved": bpc * value / world / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": bpc * value / world / 1e9 / peak,
                             "traffic": None, "note": "per-GPU EFFECTIVE bandwidth: algorithmic bytes per cell-step / time; "
                                                      "see the N=1 line for the dominant kernel"},
                "cpu_baseline": None, "impl": "b200", "kernels_launches_ms_per_step_rank0": kernels_ms,
                "per_rank_ms_per_step_gpu_and_host_enqueue": per_rank,
                "per_rank_ms_per_step_compute_and_comm_kernels": per_rank_split}
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def slab_parity(wl, fb, dist, rank, world, steps=3):
    """Every rank steps its slab of the seeded scene `steps` times; rank 0 also steps the whole global grid on its
    own GPU (the single-GPU path, itself bit-exact against the oracle in tests/) and compares the SHA-256 of every
    rank's owned rows of u, v, u_prev, v_prev, dens, dens_prev.  Returns "ok" or a description of what differs."""
    import torch
    from fruid_b200 import _lib
    from fruid_b200.dist import SlabDensitySim, SlabFluidSim
    L = _lib.lib
    w, h = wl.w, wl.h
    sim, den = SlabFluidSim(w, h, rank, world), SlabDensitySim(w, h, rank, world)
    for o in (sim, den):
        o.set_iterations(wl.iters)
    for f, name in enumerate(("u", "v", "u_prev", "v_prev")):
        sim.upload(f, scene_field(name, w, h, sim.y0, sim.y1, cfl=wl.cfl))
    den.upload(0, scene_field("dens", w, h, sim.y0, sim.y1))
    den.upload(1, scene_field("dens_prev", w, h, sim.y0, sim.y1))
    dist.barrier()
    for _ in range(steps):
        sim.step(DT, wl.visc)
        den.step(sim, DT, wl.diff)
    sim.sync(); den.sync()
    mine = [hashlib.sha256(sim.download(f).tobytes()).hexdigest() for f in range(4)]
    mine += [hashlib.sha256(den.download(f).tobytes()).hexdigest() for f in range(2)]
    y0 = sim.y0
    gathered = [None] * world
    dist.all_gather_object(gathered, (y0, sim.y1, mine))
    dist.barrier()
    del sim, den
    result = "ok"
    if rank == 0:
        try:
            s1, d1 = fb.FluidSim(w, h), fb.DensitySim(w, h)
            for o in (s1, d1):
                o.set_iterations(wl.iters)
            for f, name in enumerate(("u", "v", "u_prev", "v_prev")):
                _lib.check(L.fruid_fluid_upload(s1._h, f, scene_field(name, w, h, cfl=wl.cfl)))
            _lib.check(L.fruid_density_upload(d1._h, 0, scene_field("dens", w, h)))
            _lib.check(L.fruid_density_upload(d1._h, 1, scene_field("dens_prev", w, h)))
            for _ in range(steps):
                _lib.check(L.fruid_fluid_step(s1._h, DT, wl.visc))
                _lib.check(L.fruid_density_step_dev(d1._h, s1._h, DT, wl.diff))
            bad = []
            a = np.empty((h, w), np.float32)
            for k, name in enumerate(FIELDS6):
                if k < 4:
                    _lib.check(L.fruid_fluid_download(s1._h, k, _lib.ptr(a)))
                else:
                    _lib.check(L.fruid_density_download(d1._h, k - 4, _lib.ptr(a)))
                for r, (ya, yb, digests) in enumerate(gathered):
                    if hashlib.sha256(np.ascontiguousarray(a[ya:yb]).tobytes()).hexdigest() != digests[k]:
                        bad.append(f"{name}@rank{r}")
            result = "ok" if not bad else "MISMATCH: " + ",".join(bad)
            del s1, d1
        except Exception as e:  # e.g. out of memory for the single-GPU twin: say so instead of guessing
            result = f"not checked: {type(e).__name__}: {e}"
    box = [result]
    dist.broadcast_object_list(box, 0)
    del torch
    return f"{box[0]} ({steps} steps of the {w}x{h} grid on {world} GPUs vs 1 GPU, SHA-256 of every rank's rows of 6 fields)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 100 (1000 for inject1024, 5 per K for ksweep)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="headline",
                    choices=["headline", "cfl64", "viscous", "scene250", "scene128", "inject1024", "ksweep", "strong16k", "weak16k"])
    ap.add_argument("--size", type=int, default=0, help="grid width (default 4096)")
    ap.add_argument("--iters", type=int, default=0, help="Jacobi sweeps per lin_solve (default 20, src/lib.rs:52)")
    ap.add_argument("--dyes", type=int, default=None, help="DensitySims per scene step (default 1; 4 for the scene configs)")
    ap.add_argument("--fused", type=int, default=None, help="Jacobi sweeps per launch (default: library choice)")
    ap.add_argument("--rows", type=int, default=0, help="global rows (default: size, or size*N = weak scaling)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-variants", action="store_true", help="skip the viscous / CFL~64 variants of the headline line")
    ap.add_argument("--no-e2e", action="store_true", help="multi-GPU: skip the end-to-end leg")
    ap.add_argument("--no-parity", action="store_true", help="multi-GPU: skip the comparison with a single-GPU run")
    ap.add_argument("--no-omp", action="store_true", help="reference arm: skip the all-cores context figure")
    ap.add_argument("--sustain-seconds", type=float, default=2.0, help="also time the step loop for about this long (0 = off)")
    ap.add_argument("--ref-seconds", type=float, default=150.0, help="reference arm: CPU time budget that sizes the sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_b200(args)


if __name__ == "__main__":
    main()
