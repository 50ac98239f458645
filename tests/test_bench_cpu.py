"""CPU-only: the bench's reference arm (the CPU oracle timed as the reference's implementation of the path)
prints one well-formed JSON line, and the synthetic scene generator is deterministic and slab-consistent."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "192",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "cell-steps/sec" and line["unit"] == "cell-steps/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["dtype"] == "f32" and "workload" in line["config"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "128",
                        "--steps", "1", "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_scene_generator_is_deterministic_and_slab_consistent():
    sys.path.insert(0, ROOT)
    import bench
    full = bench.make_scene(96, 80)
    again = bench.make_scene(96, 80)
    part = bench.make_scene(96, 80, 30, 55)
    for k in ("u", "v", "u_prev", "v_prev", "dens", "dens_prev"):
        assert np.array_equal(full[k], again[k])
        assert np.array_equal(full[k][30:55], part[k]), k   # any slab of the scene can be generated on its own
    assert bench.bytes_per_cell_step(20) == 1388            # 188 + 60K
