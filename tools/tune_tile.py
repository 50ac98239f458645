"""Times the fused Jacobi kernel (fruid_dev_lin_solve) over tile shapes / fused depths on one
GPU.  Device buffers come from torch (plumbing); timing is CUDA events on the launch stream.
    python tools/tune_tile.py [--size 4096] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fruid_b200 as fb  # noqa: E402
from fruid_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--cfgs", default="8x16,10x16,6x16,6x20,7x20,5x24,4x24,8x12,4x16")
ap.add_argument("--fused", default="1,2,4,6,10,20")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--modes", default="poisson,general,inviscid")
ap.add_argument("--scale", type=float, default=1.0, help="magnitude of x and x0 (1e-35: below the range of the fused Poisson form)")
args = ap.parse_args()

w = h = args.size
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(h, w, device=dev, generator=g) * args.scale
x0 = torch.randn(h, w, device=dev, generator=g) * args.scale
s = torch.zeros(h, w, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream()
L = _lib.lib
rows = []
for (name, a, c) in (("poisson a=1 c=4", 1.0, 4.0), ("general a=.17 c=1.67", 0.1676, 1.6704), ("inviscid a=0 c=1", 0.0, 1.0)):
    if name.split()[0] not in args.modes.split(","):
        continue
    for cfg in args.cfgs.split(","):
        R, NW = map(int, cfg.split("x"))
        _lib.check(L.fruid_set_tile_config(R, NW, 0))
        for T in map(int, args.fused.split(",")):
            if T != 1 and cfg != args.cfgs.split(",")[0] and False:
                continue
            try:
                _lib.check(L.fruid_dev_lin_solve(0, x.data_ptr(), x0.data_ptr(), s.data_ptr(), w, h, w, a, c, args.iters, T, st.cuda_stream))
            except fb.FruidError as e:
                continue
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(args.reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                _lib.check(L.fruid_dev_lin_solve(0, x.data_ptr(), x0.data_ptr(), s.data_ptr(), w, h, w, a, c, args.iters, T, st.cuda_stream))
                e1.record(st)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            gbs = 12.0 * w * h * args.iters / (best * 1e-3) / 1e9
            rows.append({"mode": name, "cfg": cfg, "T": T, "ms": best, "us_per_sweep": best * 1e3 / args.iters, "eff_GBs": gbs})
            print(f"{name:22s} R x NW={cfg:6s} T={T:3d}  {best:8.3f} ms/lin_solve  {best*1e3/args.iters:7.2f} us/sweep  eff {gbs:8.0f} GB/s", flush=True)
            if T == 1:
                break_cfg = True
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/tune_tile.json", "w"), indent=1)
