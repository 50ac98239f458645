"""Host-side tiling arithmetic of the fused Jacobi kernel (fruid_debug_tile_plan: no CUDA call, runs without a GPU).

The kernel advances a 128 x (R*NW) tile by T sweeps and stores only the cells at least H = T rounded up to even away from
every tile edge that is not the domain edge; tiles therefore step by 128 - 2H / R*NW - 2H.  These tests pin the
properties the parity of every cell depends on: the tiles cover the field, nothing is skipped, the fused depth fits the
shape, and the automatic choice keeps the measured policy (small tiles only while every tile can have its own SM)."""
import ctypes

import numpy as np
import pytest

SHAPES = {(8, 16), (10, 16), (6, 16), (6, 20), (7, 20), (5, 24), (4, 24), (8, 12), (4, 16), (2, 32), (2, 16), (4, 8), (3, 16)}


@pytest.fixture(scope="module")
def L():
    import __graft_entry__
    __graft_entry__.build()
    from fruid_b200 import _lib
    return _lib


def plan(L, w, h, iters=20, fused=0):
    out = (ctypes.c_int * 6)()
    rc = L.lib.fruid_debug_tile_plan(w, h, iters, fused, out)
    return rc, list(out)


def covered(extent, tile, H, n):
    """Cells of [0, extent) stored by n tiles of `tile` cells stepping by tile - 2H (domain edges need no halo)."""
    stride = tile - 2 * H
    got = np.zeros(extent, bool)
    for k in range(n):
        o = k * stride
        lo = 0 if k == 0 else o + H
        hi = extent if k == n - 1 else o + tile - H
        assert 0 <= lo <= hi
        assert hi <= extent or k == n - 1
        # the last tile must still contain the domain edge
        if k == n - 1:
            assert o + tile >= extent, (extent, tile, H, n)
        got[lo:min(hi, extent)] = True
    return got


@pytest.mark.parametrize("iters", [2, 6, 20, 26, 40, 100])
def test_tiles_cover_the_field_for_many_sizes(L, iters):
    rng = np.random.default_rng(iters)
    sizes = [(3, 3), (4, 129), (129, 4), (128, 128), (250, 250), (512, 512), (1024, 1024), (4096, 4096), (16384, 2048)]
    sizes += [(int(a), int(b)) for a, b in rng.integers(3, 3000, (60, 2))]
    for w, h in sizes:
        rc, (R, NW, T, ntx, nty, launches) = plan(L, w, h, iters)
        assert rc == 0, (w, h)
        assert (R, NW) in SHAPES and 1 <= T <= iters
        H = T + (T & 1)
        assert 128 - 2 * H >= 4 and R * NW - 2 * H >= 4          # at least four stored rows / columns per tile
        assert launches == -(-iters // T) or launches * T >= iters
        assert covered(w, 128, H, ntx).all(), (w, h, R, NW, T, ntx)
        assert covered(h, R * NW, H, nty).all(), (w, h, R, NW, T, nty)
        # no tile is superfluous: with one tile fewer the field would not be covered
        if ntx > 1:
            assert (ntx - 2) * (128 - 2 * H) + 128 < w
        if nty > 1:
            assert (nty - 2) * (R * NW - 2 * H) + R * NW < h


def test_automatic_shape_policy(L):
    """Small grids take smaller tiles only while every tile can have its own SM (148); big grids the 128x128 tile
    with 10 sweeps per launch."""
    for n in (64, 128, 250, 300, 512, 768, 1024):
        rc, (R, NW, T, ntx, nty, _) = plan(L, n, n)
        assert rc == 0 and ntx * nty <= 148, (n, R, NW, T, ntx * nty)
        assert R * NW < 128 or n >= 1024 or T == 20, (n, R, NW, T)
    assert plan(L, 250, 250)[1][:3] == [2, 16, 10]
    assert plan(L, 512, 512)[1][:3] == [3, 16, 10]
    assert plan(L, 1024, 1024)[1][:3] == [4, 24, 10]
    # a few tiles per SM: the shape that quantises best over 148 SMs (measured per size, csrc/jacobi_tile.cuh)
    assert plan(L, 1536, 1536)[1][:3] == [6, 20, 10]
    assert plan(L, 2048, 2048)[1][:3] == [8, 16, 10]
    assert plan(L, 3072, 3072)[1][:3] == [10, 16, 10]
    for n in (4096, 16384):
        assert plan(L, n, n)[1][:3] == [7, 20, 10]
    # an explicit fused depth is respected (and moves to a shape that supports it)
    rc, p = plan(L, 250, 250, 20, 20)
    assert rc == 0 and p[2] == 20 and (p[0] * p[1]) // 2 - 2 >= 20
    rc, p = plan(L, 4096, 4096, 20, 4)
    assert rc == 0 and p[:3] == [7, 20, 4] and p[5] == 5
    assert plan(L, 4096, 4096, 20, 1)[1][2] == 1               # one plain kernel per sweep


def test_plan_rejects_bad_arguments(L):
    assert plan(L, 2, 2)[0] == L.ERR_BAD_SIZE                  # degenerate grids never reach the tile kernel
    assert plan(L, 100, 100, 7)[0] == L.ERR_ODD_ITERATIONS
    assert L.lib.fruid_debug_tile_plan(100, 100, 20, 0, None) == L.ERR_BAD_ARG
