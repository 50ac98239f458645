"""Oracle twins of the crate API (test infrastructure): FluidSim / DensitySim / Array2D
with the reference's semantics (src/lib.rs:210-294) computed by oracle/fruid_oracle.c, so
fruid_b200.scene.Scene can drive either implementation."""
import numpy as np

import oracle


class OArray2D:
    def __init__(self, arr):
        self.a = arr

    def width(self):
        return self.a.shape[1]

    def height(self):
        return self.a.shape[0]

    def __getitem__(self, pos):
        return self.a[pos[1], pos[0]]

    def __setitem__(self, pos, val):
        self.a[pos[1], pos[0]] = val

    def fill(self, val):
        self.a.fill(val)

    def data(self):
        return self.a.reshape(-1)

    data_mut = data

    def numpy(self):
        return self.a.copy()


class OFluidSim:
    iters = 20
    order = oracle.FAST

    def __init__(self, w, h):
        self.f = oracle.Fluid(w, h, iters=self.iters, order=self.order)

    def step(self, dt, visc):
        self.f.step(dt, visc)

    def uv(self):
        return OArray2D(self.f.u), OArray2D(self.f.v)

    def uv_mut(self):
        return OArray2D(self.f.u_prev), OArray2D(self.f.v_prev)

    def width(self):
        return self.f.w

    def height(self):
        return self.f.h


class ODensitySim:
    iters = 20
    order = oracle.FAST

    def __init__(self, w, h):
        self.d = oracle.Density(w, h, iters=self.iters, order=self.order)

    def step(self, uv, dt, diff):
        u, v = uv
        self.d.step((u.a, v.a), dt, diff)

    def density(self):
        return OArray2D(self.d.dens)

    def density_mut(self):
        return OArray2D(self.d.dens_prev)
