"""Multi-GPU y-slab decomposition: host-side plumbing (one process per GPU, torchrun).

torch.distributed is used ONLY to exchange the 64-byte CUDA IPC handles of each rank's device
allocation at start-up (and for the bench's barriers); all data-path traffic -- halo rows and
advection gathers -- moves over NVLink peer memory inside the library's own kernels
(fruid_b200/csrc/slab_comm.cuh).  Pure partition arithmetic lives in `partition()` so that it can
be tested on CPU (tests/test_slab_cpu.py, gloo world_size 2).

Teardown is collective: call `sync()` on every rank and then `torch.distributed.barrier()` before the
objects are dropped -- a neighbour may otherwise still be reading the memory a faster rank frees.
"""
from __future__ import annotations

import ctypes

import numpy as np

HALO = 22   # rows kept of each neighbour (csrc/slab_comm.cuh: SLAB_HALO)
MAX_T = 10  # fused Jacobi sweeps per launch allowed in slab mode (SLAB_MAX_T)


def partition(gh: int, world: int):
    """Owned global rows [own0[r], own0[r+1]) and local extents (row0[r], rows[r]) of every rank."""
    per = gh // world
    own0 = [r * per for r in range(world)] + [gh]
    row0, rows = [], []
    for r in range(world):
        lo = max(0, own0[r] - HALO) if world > 1 else 0
        hi = min(gh, own0[r + 1] + HALO) if world > 1 else gh
        row0.append(lo)
        rows.append(hi - lo)
    return own0, row0, rows


def _exchange_handles(export_fn, attach_fn, handle, rank: int, world: int, group=None) -> None:
    import torch
    import torch.distributed as dist
    buf = (ctypes.c_ubyte * 64)()
    export_fn(handle, ctypes.cast(buf, ctypes.c_void_p))
    mine = bytes(buf)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    for r, hb in enumerate(gathered):
        if r == rank:
            continue
        cbuf = (ctypes.c_ubyte * 64).from_buffer_copy(hb)
        attach_fn(handle, r, ctypes.cast(cbuf, ctypes.c_void_p))
    dist.barrier(group=group)
    del torch


class SlabFluidSim:
    """FluidSim (src/lib.rs:247-294) decomposed into y-slabs over the ranks of a process group."""

    def __init__(self, width: int, height: int, rank: int, world: int, group=None):
        from . import _lib
        self._lib = _lib
        self._h = _lib._vp()
        _lib.check(_lib.lib.fruid_fluid_create_slab(width, height, rank, world, ctypes.byref(self._h)))
        self.rank, self.world, self.w, self.gh = rank, world, width, height
        _exchange_handles(lambda h, b: _lib.check(_lib.lib.fruid_fluid_ipc_export(h, b)),
                          lambda h, r, b: _lib.check(_lib.lib.fruid_fluid_ipc_attach(h, r, b)),
                          self._h, rank, world, group)
        y0, y1 = _lib._sz(0), _lib._sz(0)
        _lib.check(_lib.lib.fruid_fluid_slab_rows(self._h, ctypes.byref(y0), ctypes.byref(y1)))
        self.y0, self.y1 = int(y0.value), int(y1.value)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        lib = getattr(getattr(self, "_lib", None), "lib", None)
        if h and lib is not None:
            lib.fruid_fluid_destroy(h)

    def upload(self, field: int, rows: np.ndarray) -> None:
        """rows: (y1 - y0, w) float32 = the owned rows of the global field."""
        assert rows.shape == (self.y1 - self.y0, self.w)
        self._lib.check(self._lib.lib.fruid_fluid_upload(self._h, field, self._lib.ptr(rows)))

    def download(self, field: int) -> np.ndarray:
        out = np.empty((self.y1 - self.y0, self.w), np.float32)
        self._lib.check(self._lib.lib.fruid_fluid_download(self._h, field, self._lib.ptr(out)))
        return out

    def inject(self, field: int, x: int, y: int, val: float) -> None:
        self._lib.check(self._lib.lib.fruid_fluid_inject(self._h, field, x, y, val))

    def set_iterations(self, k: int) -> None:
        self._lib.check(self._lib.lib.fruid_fluid_set_iterations(self._h, k))

    def step(self, dt: float, visc: float) -> None:
        self._lib.check(self._lib.lib.fruid_fluid_step(self._h, dt, visc))

    def sync(self) -> None:
        self._lib.check(self._lib.lib.fruid_fluid_sync(self._h))


class SlabDensitySim:
    """DensitySim (src/lib.rs:210-245) decomposed like its SlabFluidSim."""

    def __init__(self, width: int, height: int, rank: int, world: int, group=None):
        from . import _lib
        self._lib = _lib
        self._h = _lib._vp()
        _lib.check(_lib.lib.fruid_density_create_slab(width, height, rank, world, ctypes.byref(self._h)))
        self.rank, self.world, self.w, self.gh = rank, world, width, height
        _exchange_handles(lambda h, b: _lib.check(_lib.lib.fruid_density_ipc_export(h, b)),
                          lambda h, r, b: _lib.check(_lib.lib.fruid_density_ipc_attach(h, r, b)),
                          self._h, rank, world, group)
        own0, _, _ = partition(height, world)
        self.y0, self.y1 = own0[rank], own0[rank + 1]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        lib = getattr(getattr(self, "_lib", None), "lib", None)
        if h and lib is not None:
            lib.fruid_density_destroy(h)

    def upload(self, field: int, rows: np.ndarray) -> None:
        assert rows.shape == (self.y1 - self.y0, self.w)
        self._lib.check(self._lib.lib.fruid_density_upload(self._h, field, self._lib.ptr(rows)))

    def download(self, field: int) -> np.ndarray:
        out = np.empty((self.y1 - self.y0, self.w), np.float32)
        self._lib.check(self._lib.lib.fruid_density_download(self._h, field, self._lib.ptr(out)))
        return out

    def fill(self, field: int, val: float) -> None:
        """`density_mut().data_mut().fill(val)` (main.rs:105-108) on every rank's rows."""
        self._lib.check(self._lib.lib.fruid_density_fill(self._h, field, val))

    def inject(self, field: int, x: int, y: int, val: float) -> None:
        """One-cell write at GLOBAL (x, y); every rank calls it, the owner of the row performs it."""
        self._lib.check(self._lib.lib.fruid_density_inject(self._h, field, x, y, val))

    def set_iterations(self, k: int) -> None:
        self._lib.check(self._lib.lib.fruid_density_set_iterations(self._h, k))

    def step(self, vel: SlabFluidSim, dt: float, diff: float) -> None:
        self._lib.check(self._lib.lib.fruid_density_step_dev(self._h, vel._h, dt, diff))

    def sync(self) -> None:
        self._lib.check(self._lib.lib.fruid_density_sync(self._h))
