"""CPU-only (gloo, world_size 2): the multi-GPU y-slab scheme -- partition arithmetic, halo depth,
exchange points -- restated with the CPU oracle as the per-rank operator and gloo send/recv as the
halo exchange, must reproduce the undecomposed oracle run BIT FOR BIT on every rank's owned rows.

This mirrors op_vel_step_slab / op_dens_step_slab of fruid_b200/csrc/slab_step.cuh: exchange the
halos, run the operators on all local rows (results near a slab cut go stale one row per sweep /
stencil application, which the 22-row halo absorbs for two blocks of 10 sweeps + divergence +
gradient), advect from the owners' rows.  The GPU path itself is checked against the single-GPU
run by tools/check_multi.py on 2+ GPUs."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dist_module():
    # fruid_b200/dist.py without importing the package (whose __init__ loads the CUDA library)
    spec = importlib.util.spec_from_file_location("_fruid_dist", os.path.join(ROOT, "fruid_b200", "dist.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_partition_covers_the_grid():
    D = _dist_module()
    for gh, world in ((4096, 1), (4096, 2), (4100, 3), (16384, 8), (200, 4)):
        own0, row0, rows = D.partition(gh, world)
        assert own0[0] == 0 and own0[-1] == gh and all(b > a for a, b in zip(own0, own0[1:]))
        for r in range(world):
            lo, hi = row0[r], row0[r] + rows[r]
            assert lo == (max(0, own0[r] - D.HALO) if world > 1 else 0)
            assert hi == (min(gh, own0[r + 1] + D.HALO) if world > 1 else gh)
            if world > 1:  # halo rows come from the immediate neighbour only
                assert own0[r + 1] - own0[r] >= D.HALO
    assert D.HALO >= 2 * D.MAX_T + 2  # two fused blocks + divergence/gradient without a mid-solve exchange


def _exchange(dist, rank, world, arrs, own0, row0, rows):
    """Pull HALO rows of each field from the neighbours' owned rows (gloo send/recv)."""
    import torch
    me_lo = own0[rank] - row0[rank]
    me_hi = own0[rank + 1] - row0[rank]
    for a in arrs:
        for nbr, send_rows, recv_rows in (
            (rank - 1, slice(me_lo, me_lo + (own0[rank] - row0[rank]) if rank > 0 else me_lo), slice(0, me_lo)),
            (rank + 1, slice(me_hi - (row0[rank] + rows[rank] - own0[rank + 1]), me_hi), slice(me_hi, rows[rank])),
        ):
            if nbr < 0 or nbr >= world:
                continue
            # what the neighbour needs from me: the rows of mine that fall into ITS halo
            if nbr < rank:
                need_lo, need_hi = own0[rank], min(own0[rank + 1], row0[nbr] + rows[nbr])
            else:
                need_lo, need_hi = max(own0[rank], row0[nbr]), own0[rank + 1]
            out = torch.from_numpy(np.ascontiguousarray(a[need_lo - row0[rank]:need_hi - row0[rank]]))
            buf = torch.empty((recv_rows.stop - recv_rows.start, a.shape[1]), dtype=torch.float32)
            if rank < nbr:
                dist.send(out, nbr); dist.recv(buf, nbr)
            else:
                dist.recv(buf, nbr); dist.send(out, nbr)
            a[recv_rows] = buf.numpy()
        del send_rows


def _allgather_rows(dist, rank, world, a, own0, row0, gh):
    import torch
    parts = [torch.empty((own0[r + 1] - own0[r], a.shape[1]), dtype=torch.float32) for r in range(world)]
    mine = torch.from_numpy(np.ascontiguousarray(a[own0[rank] - row0[rank]:own0[rank + 1] - row0[rank]]))
    dist.all_gather(parts, mine) if len({p.shape for p in parts}) == 1 else _uneven_gather(dist, rank, world, parts, mine)
    return np.concatenate([p.numpy() for p in parts], axis=0)


def _uneven_gather(dist, rank, world, parts, mine):
    for r in range(world):
        if r == rank:
            parts[r].copy_(mine)
        dist.broadcast(parts[r], r)


def _worker(rank, world, port, w, gh, steps, visc, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    D = _dist_module()
    L = oracle.lib()
    own0, row0, rows = D.partition(gh, world)
    T, K, dt = D.MAX_T, 20, 0.05
    rng = np.random.default_rng(7)
    full = {k: (rng.standard_normal((gh, w)) * (3.0 if k in "uv" else 1.0)).astype(np.float32)
            for k in ("u", "v", "u0", "v0", "d", "d0")}
    # reference: undecomposed oracle
    ref = {k: a.copy() for k, a in full.items()}
    sg = np.zeros((gh, w), np.float32)
    for _ in range(steps):
        L.fruid_oracle_vel_step(ref["u"], ref["v"], ref["u0"], ref["v0"], sg, w, gh, visc, dt, K, 1)
        L.fruid_oracle_dens_step(ref["d"], ref["d0"], ref["u"], ref["v"], sg, w, gh, visc, dt, K, 1)
    # slab run
    lo, hl = row0[rank], rows[rank]
    loc = {k: np.ascontiguousarray(a[lo:lo + hl]) for k, a in full.items()}
    s = np.zeros((hl, w), np.float32)
    nx, ny = w - 2, gh - 2
    F = np.float32

    def lin_blocks(b, x, x0, a, c):
        for _ in range(K // T):  # T sweeps per block on the LOCAL array; no exchange in between (HALO >= 2T + 2)
            L.fruid_oracle_lin_solve(b, x, x0, s, w, hl, a, c, T, 1)

    def project(u, v, p, div):
        # the oracle's project uses the LOCAL height for its scale factors; restate it with the global ny
        div[1:-1, 1:-1] = 0
        p[...] = 0
        L.fruid_oracle_diff_acc(div, u, w, hl, F(-0.5) / F(nx), 0, 1)
        L.fruid_oracle_diff_acc(div, v, w, hl, F(-0.5) / F(ny), 1, 1)
        L.fruid_oracle_set_bnd(0, div, w, hl)
        lin_blocks(0, p, div, F(1.0), F(4.0))
        L.fruid_oracle_diff_acc(u, p, w, hl, F(-0.5) * F(nx), 0, 1)
        L.fruid_oracle_diff_acc(v, p, w, hl, F(-0.5) * F(ny), 1, 1)
        L.fruid_oracle_set_bnd(1, u, w, hl)
        L.fruid_oracle_set_bnd(2, v, w, hl)

    def advect_global(b, name_d, name_d0, u_name, v_name, src):
        g = {k: _allgather_rows(dist, rank, world, src[k], own0, row0, gh) for k in {name_d0, u_name, v_name}}
        out = np.zeros((gh, w), np.float32)
        L.fruid_oracle_advect(b, out, g[name_d0], g[u_name], g[v_name], w, gh, dt, 1)
        src[name_d][...] = out[lo:lo + hl]  # owned rows (and whatever else) from the owners' data

    for _ in range(steps):
        a = F(F(F(F(dt) * F(visc)) * F(nx)) * F(ny))
        c = F(1.0) + F(4.0) * a
        _exchange(dist, rank, world, [loc["u"], loc["v"], loc["u0"], loc["v0"]], own0, row0, rows)
        L.fruid_oracle_add_source(loc["u"], loc["u0"], w, hl, dt)
        L.fruid_oracle_add_source(loc["v"], loc["v0"], w, hl, dt)
        lin_blocks(1, loc["u0"], loc["u"], a, c)
        lin_blocks(2, loc["v0"], loc["v"], a, c)
        _exchange(dist, rank, world, [loc["u0"], loc["v0"]], own0, row0, rows)
        project(loc["u0"], loc["v0"], loc["u"], loc["v"])
        uv = {"u": loc["u"], "v": loc["v"], "u0": loc["u0"], "v0": loc["v0"]}
        g0 = {k: _allgather_rows(dist, rank, world, uv[k], own0, row0, gh) for k in ("u0", "v0")}
        for b, dn, d0n in ((1, "u", "u0"), (2, "v", "v0")):
            out = np.zeros((gh, w), np.float32)
            L.fruid_oracle_advect(b, out, g0[d0n], g0["u0"], g0["v0"], w, gh, dt, 1)
            loc[dn][...] = out[lo:lo + hl]
        _exchange(dist, rank, world, [loc["u"], loc["v"]], own0, row0, rows)
        project(loc["u"], loc["v"], loc["u0"], loc["v0"])
        # density
        ad = F(F(F(F(dt) * F(visc)) * F(nx)) * F(ny))
        _exchange(dist, rank, world, [loc["d"], loc["d0"]], own0, row0, rows)
        L.fruid_oracle_add_source(loc["d"], loc["d0"], w, hl, dt)
        lin_blocks(0, loc["d0"], loc["d"], ad, F(1.0) + F(4.0) * ad)
        gd = {k: _allgather_rows(dist, rank, world, loc[k], own0, row0, gh) for k in ("d0", "u", "v")}
        out = np.zeros((gh, w), np.float32)
        L.fruid_oracle_advect(0, out, gd["d0"], gd["u"], gd["v"], w, gh, dt, 1)
        loc["d"][...] = out[lo:lo + hl]
    bad = {}
    o0, o1 = own0[rank] - lo, own0[rank + 1] - lo
    for k in ("u", "v", "u0", "v0", "d", "d0"):
        got, want = loc[k][o0:o1], ref[k][own0[rank]:own0[rank + 1]]
        n = int(np.count_nonzero(got.view(np.uint32) != want.view(np.uint32)))
        if n:
            bad[k] = n
    q.put((rank, bad))
    dist.destroy_process_group()


@pytest.mark.parametrize("visc", [0.0, 3e-4])
def test_slab_scheme_matches_undecomposed_oracle_world2(visc):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + (1 if visc else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, 124, 2, visc, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, bad in res:
        assert not bad, f"rank {rank}: mismatching cells {bad}"
