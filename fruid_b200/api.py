"""Python mirror of the reference crate API (src/lib.rs:1, 210-294) over the C ABI.

Same names, argument order and semantics as the Rust types:

    FluidSim.new(w, h) / FluidSim(w, h); .step(dt, visc); .uv(); .uv_mut(); .width(); .height()
    DensitySim(w, h); .step((u, v), dt, diff); .density(); .density_mut()

The fields live in HBM; the Array2D objects handed out are host mirrors that are
downloaded lazily on first read and uploaded lazily before the next step when written.
`DensitySim.step(sim.uv(), ...)` (src/main.rs:115) recognises the mirrors of a live
FluidSim and uses its device-resident velocity instead of uploading it again.
"""
from __future__ import annotations

import mmap
import weakref

import numpy as np

from . import _lib
from ._lib import check, lib, ptr


PIN_MIN_BYTES = 64 << 10


class Array2D:
    """Host mirror of one device field: idek_basics::Array2D<f32> (src/lib.rs:1).

    Indexed by (x, y) like the Rust type; flat layout x + y*width.  Reads download the
    field if the host copy is stale; single-cell writes are forwarded as injections;
    data_mut() marks the whole field dirty (re-uploaded before the next step).
    """

    def __init__(self, owner, field: int, width: int, height: int):
        # weak: the simulation owns its mirrors, not the other way round -- no reference cycle, so dropping the
        # last reference to a FluidSim / DensitySim releases its device memory right away (like Rust's Drop)
        self._owner_ref, self._field = weakref.ref(owner), field
        self._w, self._h = width, height
        # The mirror owns whole pages (an anonymous mapping, zero-filled like Array2D::new): page-locking it can
        # then never pin, or later unpin, memory that belongs to another object.  Mirrors below PIN_MIN_BYTES are
        # not pinned at all -- their transfers are latency-bound either way.
        nbytes = width * height * 4
        self._map = mmap.mmap(-1, max(nbytes, 1))
        self._host = np.frombuffer(self._map, np.float32, width * height).reshape(height, width)
        self._host_ptr = self._host.ctypes.data
        self._pinned = nbytes >= PIN_MIN_BYTES and lib.fruid_host_register(self._host_ptr, len(self._map)) == 0
        self._stale = False   # device newer than host
        self._dirty = False   # host newer than device (whole field)
        self._fill = None     # pending whole-field fill value (performed on the device, nothing uploaded)
        self._cells = []      # pending single-cell writes (x, y, value), applied after the fill

    def _release(self):
        """Unpin; called by the owning simulation BEFORE it destroys its handle (and again, harmlessly, by __del__)."""
        if getattr(self, "_pinned", False) and lib is not None:
            self._pinned = False
            try:
                lib.fruid_host_unregister(self._host_ptr)
            except Exception:  # interpreter shutdown
                pass

    def __del__(self):
        self._release()

    @property
    def _owner(self):
        owner = self._owner_ref()
        if owner is None:
            raise ReferenceError("the simulation this Array2D belongs to has been dropped")
        return owner

    def width(self) -> int:
        return self._w

    def height(self) -> int:
        return self._h

    # -- synchronisation ---------------------------------------------------------
    def _pull(self):
        if self._stale:
            self._owner._download(self._field, self._host)
            self._stale = False
            for (x, y, val) in self._cells:  # keep not-yet-pushed writes visible
                self._host[y, x] = val

    def _push(self):
        """Make the device copy current (called by the owner before a step)."""
        if self._dirty:
            self._owner._upload(self._field, self._host)
        else:
            if self._fill is not None:
                self._owner._fill_dev(self._field, self._fill)
            for (x, y, val) in self._cells:
                self._owner._inject(self._field, x, y, val)
        self._dirty = False
        self._fill = None
        self._cells = []

    def _mark_stale(self):
        self._stale, self._dirty, self._fill, self._cells = True, False, None, []

    # -- Rust-like surface --------------------------------------------------------
    def data(self) -> np.ndarray:
        """`data()`: flat read-only view, x fastest."""
        self._pull()
        v = self._host.reshape(-1).view()
        v.flags.writeable = False
        return v

    def data_mut(self) -> np.ndarray:
        """`data_mut()`: flat writable view; the field is re-uploaded before the next step."""
        self._pull()
        self._dirty = True
        return self._host.reshape(-1)

    def fill(self, val: float) -> None:
        """`data_mut().fill(val)` (src/main.rs:105) without a download and without an upload: the device fills."""
        self._host.fill(val)
        self._stale, self._dirty, self._fill, self._cells = False, False, float(np.float32(val)), []

    def copy_from(self, src) -> None:
        """`data_mut().copy_from_slice(src)`: overwrite the whole field (no download first)."""
        np.copyto(self._host, np.asarray(src, np.float32).reshape(self._h, self._w))
        self._stale, self._dirty, self._fill, self._cells = False, True, None, []

    def __getitem__(self, pos):
        x, y = pos
        self._pull()
        return self._host[y, x]

    def __setitem__(self, pos, val):
        x, y = pos
        if not (0 <= x < self._w and 0 <= y < self._h):
            raise IndexError(f"({x}, {y}) outside {self._w}x{self._h}")  # Rust panics
        val = np.float32(val)
        if not self._stale:
            self._host[y, x] = val
        if not self._dirty:
            self._cells.append((int(x), int(y), float(val)))

    def numpy(self) -> np.ndarray:
        """(height, width) float32 copy of the current contents."""
        self._pull()
        return self._host.copy()


class FluidSim:
    """src/lib.rs:247-294."""

    def __init__(self, width: int, height: int):
        self._h = _lib._vp()
        check(lib.fruid_fluid_create(width, height, _lib.ctypes.byref(self._h)))
        self._w_, self._h_ = int(width), int(height)
        self._u = Array2D(self, _lib.U, width, height)
        self._v = Array2D(self, _lib.V, width, height)
        self._u_prev = Array2D(self, _lib.U_PREV, width, height)
        self._v_prev = Array2D(self, _lib.V_PREV, width, height)

    new = classmethod(lambda cls, width, height: cls(width, height))  # FluidSim::new, lib.rs:256

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.fruid_fluid_sync(h)  # nothing may still be reading or writing a mirror
            for a in (getattr(self, n, None) for n in ("_u", "_v", "_u_prev", "_v_prev")):
                if a is not None:
                    a._release()
            lib.fruid_fluid_destroy(h)

    # transfers used by the mirrors
    def _download(self, field, host):
        check(lib.fruid_fluid_download(self._h, field, ptr(host)))

    def _upload(self, field, host):
        check(lib.fruid_fluid_upload(self._h, field, ptr(host)))

    def _inject(self, field, x, y, val):
        check(lib.fruid_fluid_inject(self._h, field, x, y, val))

    def _fill_dev(self, field, val):
        check(lib.fruid_fluid_fill(self._h, field, val))

    def set_iterations(self, k: int) -> None:
        """Sweeps per lin_solve (reference: 20, src/lib.rs:52).  Extension."""
        check(lib.fruid_fluid_set_iterations(self._h, k))

    def set_fused_sweeps(self, t: int) -> None:
        check(lib.fruid_fluid_set_fused_sweeps(self._h, t))

    def step(self, dt: float, visc: float) -> None:  # lib.rs:267
        for a in (self._u, self._v, self._u_prev, self._v_prev):
            a._push()
        check(lib.fruid_fluid_step(self._h, dt, visc))
        for a in (self._u, self._v, self._u_prev, self._v_prev):
            a._mark_stale()

    def uv(self):  # lib.rs:279
        return (self._u, self._v)

    def uv_mut(self):  # lib.rs:283 -- (u_prev, v_prev): sources in, pressure/divergence out
        return (self._u_prev, self._v_prev)

    def width(self) -> int:  # lib.rs:287
        return self._w_

    def height(self) -> int:  # lib.rs:291
        return self._h_

    def sync(self) -> None:
        check(lib.fruid_fluid_sync(self._h))


class DensitySim:
    """src/lib.rs:210-245."""

    def __init__(self, width: int, height: int):
        self._h = _lib._vp()
        check(lib.fruid_density_create(width, height, _lib.ctypes.byref(self._h)))
        self._w_, self._h_ = int(width), int(height)
        self._dens = Array2D(self, _lib.DENS, width, height)
        self._dens_prev = Array2D(self, _lib.DENS_PREV, width, height)

    new = classmethod(lambda cls, width, height: cls(width, height))  # lib.rs:217

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.fruid_density_sync(h)
            for a in (getattr(self, n, None) for n in ("_dens", "_dens_prev")):
                if a is not None:
                    a._release()
            lib.fruid_density_destroy(h)

    def _download(self, field, host):
        check(lib.fruid_density_download(self._h, field, ptr(host)))

    def _upload(self, field, host):
        check(lib.fruid_density_upload(self._h, field, ptr(host)))

    def _inject(self, field, x, y, val):
        check(lib.fruid_density_inject(self._h, field, x, y, val))

    def _fill_dev(self, field, val):
        check(lib.fruid_density_fill(self._h, field, val))

    def set_iterations(self, k: int) -> None:
        check(lib.fruid_density_set_iterations(self._h, k))

    def set_fused_sweeps(self, t: int) -> None:
        check(lib.fruid_density_set_fused_sweeps(self._h, t))

    def step(self, uv, dt: float, diff: float) -> None:  # lib.rs:226 (uv first, then dt, diff)
        u, v = uv
        self._dens._push()
        self._dens_prev._push()
        sim = u._owner_ref() if isinstance(u, Array2D) else None
        if (isinstance(u, Array2D) and isinstance(v, Array2D) and isinstance(sim, FluidSim)
                and v._owner_ref() is sim and u._field == _lib.U and v._field == _lib.V):
            u._push()
            v._push()
            check(lib.fruid_density_step_dev(self._h, sim._h, dt, diff))
        else:
            un = np.ascontiguousarray(u.numpy() if isinstance(u, Array2D) else u, np.float32)
            vn = np.ascontiguousarray(v.numpy() if isinstance(v, Array2D) else v, np.float32)
            if un.shape != (self._h_, self._w_) or vn.shape != (self._h_, self._w_):
                raise _lib.FruidError(_lib.ERR_SIZE_MISMATCH, "velocity shape does not match the density grid")
            check(lib.fruid_density_step_host(self._h, ptr(un), ptr(vn), dt, diff))
        self._dens._mark_stale()
        self._dens_prev._mark_stale()

    @staticmethod
    def step_all(dyes, uv, dt: float, diff: float) -> None:
        """`for d in dyes: d.step(uv, dt, diff)` (main.rs:114-118) as one batched call when `uv` is a
        FluidSim's device-resident uv(); same results bit for bit."""
        dyes = list(dyes)
        u, v = uv
        sim = u._owner_ref() if isinstance(u, Array2D) else None
        if not (len(dyes) > 1 and isinstance(u, Array2D) and isinstance(v, Array2D) and isinstance(sim, FluidSim)
                and v._owner_ref() is sim and u._field == _lib.U and v._field == _lib.V):
            for d in dyes:
                d.step(uv, dt, diff)
            return
        for d in dyes:
            d._dens._push()
            d._dens_prev._push()
        u._push()
        v._push()
        handles = (_lib._vp * len(dyes))(*[d._h for d in dyes])
        check(lib.fruid_density_step_dev_batch(handles, len(dyes), sim._h, dt, diff))
        for d in dyes:
            d._dens._mark_stale()
            d._dens_prev._mark_stale()

    def density(self) -> Array2D:  # lib.rs:238
        return self._dens

    def density_mut(self) -> Array2D:  # lib.rs:242 -- dens_prev (the source), not dens
        return self._dens_prev

    def sync(self) -> None:
        check(lib.fruid_density_sync(self._h))
