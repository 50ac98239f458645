"""Latency of the slab primitives (halo exchange, all-rank barrier) without any compute.
    torchrun --nproc-per-node N tools/slab_latency.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fruid_b200 import _lib
from fruid_b200.dist import SlabFluidSim
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sim = SlabFluidSim(w, 4096 * world, rank, world)
L = _lib.lib
for what, fn in (("exchange x1", lambda n: L.fruid_fluid_debug_exchange(sim._h, 1, n)),
                 ("exchange x2", lambda n: L.fruid_fluid_debug_exchange(sim._h, 2, n)),
                 ("exchange x4", lambda n: L.fruid_fluid_debug_exchange(sim._h, 4, n)),
                 ("barrier", lambda n: L.fruid_fluid_debug_barrier(sim._h, n))):
    _lib.check(fn(20)); sim.sync(); dist.barrier()
    t0 = time.perf_counter(); _lib.check(fn(200)); sim.sync(); dt = time.perf_counter() - t0
    if rank == 0:
        print(f"{what:12s}: {dt / 200 * 1e6:7.1f} us per event (world {world}, w {w})", flush=True)
    dist.barrier()
dist.destroy_process_group()
