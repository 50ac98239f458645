#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (cell-steps/sec of the stable-fluids scene step) on B200.

A "step" is one SCENE STEP = 1x FluidSim::step (src/lib.rs:267) + 1x DensitySim::step
(src/lib.rs:226) over one w x h grid, K=20 Jacobi sweeps per lin_solve (src/lib.rs:52), i.e.
40 pressure sweeps per step -- BASELINE.json configs[2] (4096x4096, the roofline headline).
cell-steps/s = w*h*steps / seconds (border cells included, like the constructor arguments).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size 4096]

Prints ONE JSON line on rank 0.  Keys (see the task contract): value = device-timed
throughput with all fields resident in HBM; e2e = the same metric through the crate-API
mirror with HOST buffers (h2d of the three source fields + d2h of the density every step);
roofline = the dominant kernel against the measured HBM peak; cpu_baseline = the CPU oracle
(faithful loop order, 1 thread, like the single-threaded reference) on a bounded sample.
--impl reference times that CPU path alone (the Rust reference itself cannot be built in
this image: no rustc/cargo, un-vendored deps).
"""
from __future__ import annotations

import argparse
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


def make_scene(w: int, h: int, y0: int = 0, y1: int | None = None, cfl: float = 2.0):
    """Rows [y0, y1) of the global w x h scene: dict of float32 (rows, w) arrays."""
    y1 = h if y1 is None else y1
    amp = np.float32(cfl / (DT * (w - 2)))
    xh = (np.arange(w, dtype=np.float32) / np.float32(w))[None, :]
    yh = (np.arange(y0, y1, dtype=np.float32) / np.float32(h))[:, None]
    two_pi = np.float32(2 * np.pi)
    u = amp * np.sin(two_pi * xh) * np.cos(two_pi * yh) + np.float32(1e-3) * noise(0, w, y0, y1)
    v = -amp * np.cos(two_pi * xh) * np.sin(two_pi * yh) + np.float32(1e-3) * noise(1, w, y0, y1)
    sig = np.float32(w / 16.0)
    dx = np.arange(w, dtype=np.float32)[None, :] - np.float32(w / 2)
    dy = np.arange(y0, y1, dtype=np.float32)[:, None] - np.float32(h / 2)
    dens = np.exp(-(dx * dx + dy * dy) / (np.float32(2.0) * sig * sig)).astype(np.float32) \
        + np.float32(1e-3) * noise(2, w, y0, y1)
    z = np.zeros((y1 - y0, w), np.float32)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    # u_prev / v_prev / dens_prev are the per-step SOURCE fields (uv_mut / density_mut)
    return {"u": f32(u), "v": f32(v), "u_prev": f32(np.float32(1e-2) * noise(3, w, y0, y1)),
            "v_prev": f32(np.float32(1e-2) * noise(4, w, y0, y1)), "dens": f32(dens),
            "dens_prev": f32(np.float32(1e-2) * (noise(5, w, y0, y1) + np.float32(1.0))), "zero": z}


def bytes_per_cell_step(k: int = ITERS) -> int:
    """Algorithmic HBM bytes per cell per scene step: 188 + 60K."""
    return 188 + 60 * k


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
def cpu_scene_steps(w: int, rows: int, steps: int, warmup: int = 0):
    """Times `steps` scene steps of the faithful-order, single-thread oracle on a w x rows grid
    (a bounded sample of the w x w workload).  Returns (cell_steps_per_s, seconds_per_step)."""
    import oracle
    oracle.build()
    try:
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})  # pin: the reference is single-threaded
    except Exception:
        pass
    sc = make_scene(w, rows)
    f = oracle.Fluid(w, rows, iters=ITERS, order=oracle.FAITHFUL, openmp=False)
    d = oracle.Density(w, rows, iters=ITERS, order=oracle.FAITHFUL, openmp=False)
    f.u[:], f.v[:], f.u_prev[:], f.v_prev[:] = sc["u"], sc["v"], sc["u_prev"], sc["v_prev"]
    d.dens[:], d.dens_prev[:] = sc["dens"], sc["dens_prev"]
    for _ in range(warmup):
        f.step(DT, VISC)
        d.step((f.u, f.v), DT, DIFF)
    t0 = time.perf_counter()
    for _ in range(steps):
        f.step(DT, VISC)
        d.step((f.u, f.v), DT, DIFF)
    sec = (time.perf_counter() - t0) / max(steps, 1)
    return w * rows / sec, sec


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; 1 thread
    because src/lib.rs is single-threaded), same metric/config, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = args.size
    total = args.steps + args.warmup
    rows = 1024 if total <= 8 else (512 if total <= 16 else (256 if total <= 40 else 128))
    rows = min(rows, w)
    value, sec = cpu_scene_steps(w, rows, args.steps, args.warmup)
    sample = (f"each step = 1 scene step (K={ITERS}, D=1) of the faithful-order C oracle on a {w}x{rows} grid "
              f"(1/{w // rows} of the {w}x{w} workload), 1 thread pinned; the Rust reference cannot be built here")
    line = {
        "impl": "reference", "metric": "cell-steps/sec", "value": value, "unit": "cell-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 * (w / rows), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "cell-steps/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_gpus):
    w = args.size
    h = args.rows if (args.rows and n_gpus == 1) else w
    return {"workload": f"{w}x{h} scene step: FluidSim::step + 1x DensitySim::step, K={ITERS} sweeps per lin_solve "
                        f"(40 pressure sweeps/step), dt={DT}, visc={VISC}, diff={DIFF} (BASELINE.json configs[2])",
            "grid": [w, h], "jacobi_iters": ITERS, "dyes": 1, "dt": DT, "visc": VISC, "diff": DIFF,
            "init": "Taylor-Green velocity at CFL~2 cells + hash noise; Gaussian dye; noisy source fields",
            "l2": f"inputs larger than L2 (8 fields x {w * h * 4 >> 20} MiB)", "parallelism": f"slab{n_gpus}"}


# ------------------------------------------------------------------------------------------
# B200 path
# ------------------------------------------------------------------------------------------
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
    L = _lib.lib

    w = h = args.size
    if world > 1:
        return run_b200_multi(args, fb, dist, rank, world, local)
    if args.rows:
        h = args.rows  # a single GPU's share of a decomposed grid (the weak-scaling baseline of SURVEY.md 8d config 4)

    sc = make_scene(w, h)
    sim, den = fb.FluidSim(w, h), fb.DensitySim(w, h)
    if args.fused is not None:
        sim.set_fused_sweeps(args.fused)
        den.set_fused_sweeps(args.fused)
    for name, field in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
        _lib.check(L.fruid_fluid_upload(sim._h, field, _lib.ptr(sc[name])))
    for name, field in (("dens", 0), ("dens_prev", 1)):
        _lib.check(L.fruid_density_upload(den._h, field, _lib.ptr(sc[name])))
    stream = torch.cuda.ExternalStream(int(L.fruid_fluid_stream(sim._h)))

    def scene_step():
        _lib.check(L.fruid_fluid_step(sim._h, DT, VISC))
        _lib.check(L.fruid_density_step_dev(den._h, sim._h, DT, DIFF))

    # ---- device-resident throughput ("value") ----
    for _ in range(args.warmup):
        scene_step()
    torch.cuda.synchronize()
    n0 = fb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local, period_s=0.01) as clk:
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.steps):
            scene_step()
        e1.record(stream)   # density work is ordered before this by the cross-stream events
        torch.cuda.synchronize()
    launches = fb.launch_count() - n0
    ms = e0.elapsed_time(e1)
    ms_per_step = ms / args.steps
    value = w * h * args.steps / (ms * 1e-3)

    # ---- per-kernel timing pass (roofline of the dominant kernel) ----
    _lib.profile_enable(True)
    for _ in range(max(2, min(args.steps, 5))):
        scene_step()
    stats = _lib.profile_report()
    _lib.profile_enable(False)
    tot_ms = sum(t for _, t in stats.values()) or 1.0
    jac = {k: v for k, v in stats.items() if k.startswith("k_jacobi")}
    dom_name = max(jac, key=lambda k: jac[k][1]) if jac else max(stats, key=lambda k: stats[k][1])
    dom_launches, dom_ms = stats[dom_name]
    sweeps_per_scene_step = 5 * ITERS
    scene_steps_profiled = max(2, min(args.steps, 5))
    sweeps_per_launch = sweeps_per_scene_step * scene_steps_profiled / dom_launches
    alg_bytes_per_launch = 12.0 * w * h * sweeps_per_launch      # 12 B/cell/sweep
    avg_s = dom_ms * 1e-3 / dom_launches
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes_per_launch / avg_s / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom_name)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                "sweeps_per_launch": sweeps_per_launch, "avg_launch_us": avg_s * 1e6,
                "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                "kernel_share_of_step": dom_ms / tot_ms,
                "kernels_ms_per_step": {k: round(t / scene_steps_profiled, 4) for k, (n, t) in sorted(stats.items())},
                "step_effective_gbs": bytes_per_cell_step() * value / 1e9,
                "step_effective_frac": bytes_per_cell_step() * value / 1e9 / peak}

    # ---- variant: non-zero viscosity / diffusivity (general divisor c = 1 + 4a in the three diffusion solves) ----
    def timed_steps(visc, diff, n):
        for _ in range(3):
            _lib.check(L.fruid_fluid_step(sim._h, DT, visc))
            _lib.check(L.fruid_density_step_dev(den._h, sim._h, DT, diff))
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(n):
            _lib.check(L.fruid_fluid_step(sim._h, DT, visc))
            _lib.check(L.fruid_density_step_dev(den._h, sim._h, DT, diff))
        a1.record(stream)
        torch.cuda.synchronize()
        return a0.elapsed_time(a1) / n
    vms = timed_steps(1e-6, 1e-6, max(2, min(args.steps, 10)))
    variant = {"visc": 1e-6, "diff": 1e-6, "ms_per_step": vms, "value": w * h / (vms * 1e-3), "unit": "cell-steps/s",
               "note": "same scene with visc = diff = 1e-6 (SURVEY.md 8d config 1 variant B): the diffusion solves divide by "
                       "c = 1 + 4a through the verified reciprocal path"}

    # ---- end to end through the crate-API mirror with HOST buffers ----
    e2e = run_e2e(args, fb, torch, sc, w, h)

    # ---- CPU baseline beside it (bounded sample) ----
    cpu = None
    if not args.no_cpu:
        rows = min(1024, h)
        cv, csec = cpu_scene_steps(w, rows, 1)
        cpu = {"value": cv, "unit": "cell-steps/s", "cores": 1, "kind": "port",
               "sample": f"1 scene step (K={ITERS}, D=1) of the faithful-order C oracle on a {w}x{rows} grid "
                         f"({csec:.1f} s), 1 thread pinned (the reference is single-threaded); "
                         f"host has {os.cpu_count()} cores"}

    line = {"metric": "cell-steps/sec", "value": value, "unit": "cell-steps/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, 1),
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "impl": "b200", "variant_viscous": variant}
    print(json.dumps(line), flush=True)


def run_e2e(args, fb, torch, sc, w, h):
    """Scene step through the C ABI with HOST buffers: every step the three source fields (u_prev, v_prev,
    dens_prev) come from page-locked host arrays (h2d) and the resulting density goes back to a host array
    (d2h).  The transfers use the ABI's asynchronous variants (fruid_*_upload_async / _download_async): the
    sources of step i+1 travel while step i computes and the density of step i travels while step i+1
    computes.  Wall clock over n steps including every copy, synchronised on both sides; the serial
    (blocking upload / step / download) figure is reported beside it."""
    from fruid_b200 import _lib
    L = _lib.lib
    sim, den = fb.FluidSim(w, h), fb.DensitySim(w, h)
    if args.fused is not None:
        sim.set_fused_sweeps(args.fused)
        den.set_fused_sweeps(args.fused)
    for name, field in (("u", 0), ("v", 1)):
        _lib.check(L.fruid_fluid_upload(sim._h, field, _lib.ptr(sc[name])))
    _lib.check(L.fruid_density_upload(den._h, 0, _lib.ptr(sc["dens"])))
    # two sets of host source arrays and two result arrays (what a double-buffering caller keeps)
    host = [{k: np.ascontiguousarray(sc[k]).copy() for k in ("u_prev", "v_prev", "dens_prev")} for _ in range(2)]
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
            _lib.check(L.fruid_fluid_step(sim._h, DT, VISC))              # consumes the staged sources
            _lib.check(L.fruid_density_step_dev(den._h, sim._h, DT, DIFF))
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
            _lib.check(L.fruid_fluid_step(sim._h, DT, VISC))
            _lib.check(L.fruid_density_step_dev(den._h, sim._h, DT, DIFF))
            _lib.check(L.fruid_density_download(den._h, _lib.DENS, _lib.ptr(out[i & 1])))

    n = max(2, min(args.steps, 10))
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


def run_b200_multi(args, fb, dist, rank, world, local):
    """N GPUs of one node, one process per GPU: the scene is y-slab decomposed (rank r owns rows
    [r*h/N, (r+1)*h/N) plus 22 halo rows of each neighbour); halo rows and advection gathers move
    over NVLink peer memory inside the library's kernels.  WEAK scaling: w x (w*N) cells, i.e. the
    single-GPU workload per GPU.  Time = max over ranks of the CUDA-event time of the K steps,
    bracketed by barriers."""
    import torch
    from fruid_b200 import _lib
    from fruid_b200.dist import SlabDensitySim, SlabFluidSim
    L = _lib.lib
    w = args.size
    h = args.rows if args.rows else args.size * world
    sim = SlabFluidSim(w, h, rank, world)
    den = SlabDensitySim(w, h, rank, world)
    sc = make_scene(w, h, sim.y0, sim.y1)
    for name, field in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
        sim.upload(field, sc[name])
    for name, field in (("dens", 0), ("dens_prev", 1)):
        den.upload(field, sc[name])
    stream = torch.cuda.ExternalStream(int(L.fruid_fluid_stream(sim._h)))

    def scene_step():
        sim.step(DT, VISC)
        den.step(sim, DT, DIFF)

    for _ in range(args.warmup):
        scene_step()
    sim.sync(); den.sync()
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
        for _ in range(args.steps):
            scene_step()
        e1.record(stream)
        t_enq = time.perf_counter() - t_host
        sim.sync(); den.sync()
        torch.cuda.synchronize()
    dist.barrier()
    launches = fb.launch_count() - n0
    mine_ms = e0.elapsed_time(e1)
    ms = torch.tensor([mine_ms], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    per_rank = [None] * world
    dist.all_gather_object(per_rank, (round(mine_ms / args.steps, 4), round(t_enq * 1e3 / args.steps, 4)))
    value = w * h * args.steps / (ms * 1e-3)

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
    host = [{k: np.ascontiguousarray(sc[k]).copy() for k in ("u_prev", "v_prev", "dens_prev")} for _ in range(2)]
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
        sim.sync()
        den.sync()

    def serial(n):
        for i in range(n):
            hs = host[i & 1]
            sim.upload(_lib.U_PREV, hs["u_prev"])
            sim.upload(_lib.V_PREV, hs["v_prev"])
            den.upload(_lib.DENS_PREV, hs["dens_prev"])
            scene_step()
            _lib.check(L.fruid_density_download(den._h, _lib.DENS, _lib.ptr(out[i & 1])))

    n = max(2, min(args.steps, 10))
    res = {}
    for name, fn in (("pipelined", pipelined), ("serial", serial)):
        fn(2)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        fn(n)
        torch.cuda.synchronize(); dist.barrier()
        sec = torch.tensor([time.perf_counter() - t0], device="cuda")
        dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        res[name] = w * h * n / float(sec.item())
    for a in pinned:
        L.fruid_host_unregister(a.ctypes.data)
    e2e = {"value": res["pipelined"], "unit": "cell-steps/s", "h2d_bytes_per_step": 3 * w * h * 4,
           "d2h_bytes_per_step": w * h * 4, "steps": n, "serial_value": res["serial"],
           "pinned_host": len(pinned) == len(arrays),
           "note": "whole-job bytes; each rank moves its own rows (h2d of 3 source fields, d2h of density) every step; "
                   "value = asynchronous (overlapped) transfers, serial_value = blocking upload/step/download"}

    if rank == 0:
        cfg = workload_config(args, world)
        cfg["workload"] = (f"{w}x{h} scene step y-slab decomposed over {world} GPUs ({w}x{h // world} cells per GPU, weak "
                           f"scaling of the {w}x{w} single-GPU workload); FluidSim::step + 1x DensitySim::step, K={ITERS}")
        cfg["grid"] = [w, h]
        cfg["halo_rows"] = 22
        line = {"metric": "cell-steps/sec", "value": value, "unit": "cell-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak" if not args.rows else "custom", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg, "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": bytes_per_cell_step() * value / world / 1e9,
                             "peak": 6551.4, "unit": "GB/s", "frac": bytes_per_cell_step() * value / world / 1e9 / 6551.4,
                             "traffic": None, "note": "per-GPU algorithmic bytes (188+60K per cell-step) / time; see the "
                                                      "N=1 line for the dominant kernel"},
                "cpu_baseline": None, "impl": "b200", "kernels_launches_ms_per_step_rank0": kernels_ms,
                "per_rank_ms_per_step_gpu_and_host_enqueue": per_rank,
                "per_rank_ms_per_step_compute_and_comm_kernels": per_rank_split}
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--fused", type=int, default=None, help="Jacobi sweeps per launch (default: library choice)")
    ap.add_argument("--rows", type=int, default=0, help="multi-GPU: global rows (default size*N = weak scaling)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
