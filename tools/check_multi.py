"""Multi-GPU parity check: the slab-decomposed run must equal the single-GPU run BIT FOR BIT.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_multi.py

Every rank steps its slab of a seeded scene; rank 0 additionally runs the whole field on its own
GPU (single-GPU path, itself bit-exact against the oracle in tests/) and compares every rank's
owned rows of u, v, u_prev, v_prev, dens, dens_prev after `--steps` scene steps.
"""
import argparse
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (scene generator)
import fruid_b200 as fb  # noqa: E402
from fruid_b200 import _lib  # noqa: E402
from fruid_b200.dist import SlabDensitySim, SlabFluidSim  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--w", type=int, default=1000)
ap.add_argument("--h", type=int, default=1536)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--cfl", type=float, default=40.0)
ap.add_argument("--visc", type=float, default=1e-6)
ap.add_argument("--one-device", action="store_true",
                help="all ranks share cuda:0 (CUDA IPC between processes on one device; gloo instead of NCCL for the handle "
                     "exchange): lets a single-GPU box check the slab kernels, time-sliced and therefore slow")
ap.add_argument("--frame-inputs", action="store_true",
                help="between steps write sources like the reference's frame loop (main.rs:98-108): two velocity cells next "
                     "to a slab cut, dens_prev filled with 0 plus one dye cell -- host writes while other ranks may still "
                     "be gathering from the fields")
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
if args.one_device:
    local = 0
torch.cuda.set_device(local)
if args.one_device:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w, h = args.w, args.h
DT = bench.DT
sim = SlabFluidSim(w, h, rank, world)
den = SlabDensitySim(w, h, rank, world)
sc = bench.make_scene(w, h, sim.y0, sim.y1, cfl=args.cfl)
for name, field in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
    sim.upload(field, sc[name])
for name, field in (("dens", 0), ("dens_prev", 1)):
    den.upload(field, sc[name])
dist.barrier()
def frame_inputs(k, inject_u, inject_d, fill_d):
    """The per-frame source writes; the same global cells for the slab run and the single-GPU run."""
    cut = (h // world) if world > 1 else h // 2
    x = (w // 3 + 17 * k) % (w - 2) + 1
    for dy in (-1, 0):                      # the rows on either side of the first slab cut
        inject_u(2, x, cut + dy, -4500.0 + k)
        inject_u(3, x, cut + dy, 3000.0 - k)
    fill_d(1, 0.0)
    inject_d(1, x, cut - 1, 80.0 * 250 * 250)


for k in range(args.steps):
    if args.frame_inputs:
        frame_inputs(k, sim.inject, den.inject, den.fill)
    sim.step(DT, args.visc)
    den.step(sim, DT, args.visc)
sim.sync()
den.sync()
mine = [sim.download(f) for f in range(4)] + [den.download(f) for f in range(2)]
big = w * h > (1 << 23)  # large grids: compare SHA-256 digests of every rank's rows instead of gathering the fields
if big:
    mine = [hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() for a in mine]
gathered = [None] * world
dist.all_gather_object(gathered, mine)
ok = True
if rank == 0:
    full = bench.make_scene(w, h, cfl=args.cfl)
    s1, d1 = fb.FluidSim(w, h), fb.DensitySim(w, h)
    for name, field in (("u", 0), ("v", 1), ("u_prev", 2), ("v_prev", 3)):
        _lib.check(_lib.lib.fruid_fluid_upload(s1._h, field, _lib.ptr(full[name])))
    for name, field in (("dens", 0), ("dens_prev", 1)):
        _lib.check(_lib.lib.fruid_density_upload(d1._h, field, _lib.ptr(full[name])))
    for k in range(args.steps):
        if args.frame_inputs:
            frame_inputs(k, lambda f, x, y, v: _lib.check(_lib.lib.fruid_fluid_inject(s1._h, f, x, y, v)),
                         lambda f, x, y, v: _lib.check(_lib.lib.fruid_density_inject(d1._h, f, x, y, v)),
                         lambda f, v: _lib.check(_lib.lib.fruid_density_fill(d1._h, f, v)))
        _lib.check(_lib.lib.fruid_fluid_step(s1._h, DT, args.visc))
        _lib.check(_lib.lib.fruid_density_step_dev(d1._h, s1._h, DT, args.visc))
    ref = []
    for f in range(4):
        a = np.empty((h, w), np.float32)
        _lib.check(_lib.lib.fruid_fluid_download(s1._h, f, _lib.ptr(a)))
        ref.append(a)
    for f in range(2):
        a = np.empty((h, w), np.float32)
        _lib.check(_lib.lib.fruid_density_download(d1._h, f, _lib.ptr(a)))
        ref.append(a)
    names = ("u", "v", "u_prev", "v_prev", "dens", "dens_prev")
    for k, name in enumerate(names):
        if big:
            cuts = [r * (h // world) for r in range(world)] + [h]
            bad_ranks = [r for r in range(world)
                         if hashlib.sha256(np.ascontiguousarray(ref[k][cuts[r]:cuts[r + 1]]).tobytes()).hexdigest() != gathered[r][k]]
            print(f"{name:10s}: sha256 of each rank's rows vs the single-GPU run: mismatching ranks {bad_ranks}")
            ok &= not bad_ranks
            continue
        got = np.concatenate([gathered[r][k] for r in range(world)], axis=0)
        bad = int(np.count_nonzero(got.view(np.uint32) != ref[k].view(np.uint32)))
        rows = np.unique(np.nonzero(got.view(np.uint32) != ref[k].view(np.uint32))[0])[:10] if bad else []
        print(f"{name:10s}: {bad} mismatching cells of {got.size}  absmax {np.abs(ref[k]).max():.4g} {list(rows)}")
        ok &= bad == 0
    print("MULTI-GPU PARITY", "OK" if ok else "FAILED", f"world={world} grid={w}x{h} steps={args.steps} cfl={args.cfl}")
flag = torch.tensor([1 if ok else 0], device="cpu" if args.one_device else "cuda")
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
