"""The y-slab decomposition on a single-GPU box: two processes share cuda:0 and talk through CUDA IPC, so the halo
exchange kernel (`k_slab_pull`), the two-pass advection (`k_advect4<.,1>`, `k_advect4<.,2>` gathering from the owner's
memory) and the exchange bookkeeping of `lin_solve` run exactly as they do across NVLink -- only slower, because the two
contexts are time-sliced.  `tools/check_multi.py` compares every rank's owned rows of all six fields with a one-process
run of the same global grid, bit for bit (the one-process path is itself compared with the oracle in
tests/test_parity_gpu.py)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, *extra, advect_tile=None):
    env = dict(os.environ)
    if advect_tile is not None:  # pin the advect kernel of pass 1 (the library chooses by grid size otherwise)
        env["FRUID_B200_ADVECT_TILE"] = str(advect_tile)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "check_multi.py"), "--one-device", *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and "MULTI-GPU PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


# CFL 2: taps stay within the halo rows; 40: taps reach into the neighbour's owned rows; 300: taps beyond the
# neighbour (clamped at the domain edge, lib.rs:92-110).
@pytest.mark.gpu
@pytest.mark.parametrize("cfl", [2.0, 40.0, 300.0])
@pytest.mark.parametrize("advect_tile", [0, 1])
def test_two_slabs_on_one_device_match_one_process(cfl, advect_tile):
    """advect_tile: pass 1 of the advection as k_advect4<., 1> (global gathers) or k_advect_tile<., ., 1> (shared-memory
    window); pass 2 (remote taps) is k_advect4<., 2> either way and must pick up exactly the quads pass 1 left."""
    _run(2, "--w", "300", "--h", "256", "--steps", "3", "--cfl", str(cfl), "--visc", "1e-6", advect_tile=advect_tile)


@pytest.mark.gpu
def test_two_slabs_with_frame_inputs_next_to_the_cut():
    """Host writes (inject, fill) between steps while the other rank may still be gathering (main.rs:98-108)."""
    _run(2, "--w", "260", "--h", "192", "--steps", "4", "--cfl", "40", "--visc", "0", "--frame-inputs", advect_tile=1)


@pytest.mark.gpu
def test_three_slabs_on_one_device():
    """A middle rank has two cuts (two outboxes, two neighbours to wait for)."""
    _run(3, "--w", "200", "--h", "300", "--steps", "2", "--cfl", "40", "--visc", "1e-6", advect_tile=1)
