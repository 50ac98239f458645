"""A few scene steps of the bench workload (for `ncu`): python tools/ncu_step.py [size] [steps] [config]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: F401  (CUDA context like bench.py)
import bench
import fruid_b200 as fb
from fruid_b200 import _lib
size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfl = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
visc = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
if os.environ.get("FRUID_NO_GRAPHS") == "1":
    _lib.lib.fruid_set_graphs_enabled(0)
wl = bench.Workload("ncu", size, size, 20, 1, cfl=cfl, visc=visc, diff=visc)
rig = bench.Rig(fb, wl)
for _ in range(steps):
    rig.step()
_lib.check(_lib.lib.fruid_fluid_sync(rig.sim._h))
_lib.check(_lib.lib.fruid_density_sync(rig.dyes[0]._h))
print("ok", fb.launch_count())
