"""Headless restatement of the reference viewer's scene logic (src/main.rs:36-52 setup,
src/main.rs:91-118 per-frame inject / clear / step), without any rendering.

The driver is duck-typed over the crate API (FluidSim / DensitySim / Array2D), so the
same code drives the CUDA implementation and the tests' oracle twins.  cos/sin of the
injection law are evaluated once on the host in float32 and the same values are fed to
whichever implementation is being driven (src/main.rs:96,101-102 use f32::cos/sin).
"""
from __future__ import annotations

import numpy as np

F = np.float32


class Scene:
    """The default scene: one FluidSim + four dye DensitySims (c, m, y, k)."""

    def __init__(self, fluid_cls, density_cls, width: int = 250, height: int = 250, n_dyes: int = 4,
                 batched: bool = True):
        self.batched = batched
        self.sim = fluid_cls(width, height)                       # main.rs:36
        self.dyes = [density_cls(self.sim.width(), self.sim.height()) for _ in range(n_dyes)]  # main.rs:38
        self.frame_count = 0
        w, h = self.sim.width(), self.sim.height()
        intensity = F(80.0) * F(w * h)                            # main.rs:42
        # main.rs:43-46 -- order c, m, y, k; k sits at 2w/5 with intensity/100
        place = [(w // 5, intensity), (3 * w // 5, intensity), (4 * w // 5, intensity),
                 (2 * w // 5, intensity / F(100.0))]
        for d, (x, val) in zip(self.dyes, place):
            d.density_mut()[(x, h // 2)] = val
        self.sim.step(0.1, 0.0)                                   # main.rs:48
        self._step_dyes(0.1, 0.0)                                 # main.rs:49-52

    def _step_dyes(self, dt: float, diff: float) -> None:
        """`d.step(sim.uv(), dt, diff)` for every dye; batched when the implementation offers it."""
        step_all = getattr(type(self.dyes[0]), "step_all", None) if self.batched and self.dyes else None
        if step_all is not None:
            step_all(self.dyes, self.sim.uv(), dt, diff)
        else:
            for d in self.dyes:
                d.step(self.sim.uv(), dt, diff)

    def injection(self):
        """(pos, u_val, v_val) for the current frame_count, main.rs:92-102."""
        w, h = self.sim.width(), self.sim.height()
        time = F(self.frame_count) / F(12.0)
        cx, cy = w // 2, h // 2
        x = F(cx) * (np.cos(time / F(5.0), dtype=F) + F(1.0))
        # `x as usize`; x >= 0 here.  When cos() rounds to exactly 1.0 (first at frame 377: 377/60 is within 1.5e-4
        # of 2*pi) x equals the width and the reference's `u[pos] = ..` panics (index out of bounds, main.rs:100-101);
        # the headless driver keeps running with the cell clamped to the last column ("guard px < w").
        px = min(int(x), w - 1)
        u_val = F(-4500.0) * np.cos(time * F(3.0), dtype=F)
        v_val = F(-4500.0) * np.sin(time * F(3.0), dtype=F)
        return (px, cy), u_val, v_val

    def frame(self, dt: float = 1e-2, visc: float = 0.0, diff: float = 0.0) -> None:
        self.frame_count += 1                                     # main.rs:91
        pos, u_val, v_val = self.injection()
        u, v = self.sim.uv_mut()                                  # main.rs:98
        u[pos] = u_val                                            # main.rs:101
        v[pos] = v_val                                            # main.rs:102
        for d in self.dyes:                                       # main.rs:105-108
            d.density_mut().fill(0.0)
        self.sim.step(dt, visc)                                   # main.rs:114
        self._step_dyes(dt, diff)                                 # main.rs:115-118

    def run(self, frames: int, **kw) -> None:
        for _ in range(frames):
            self.frame(**kw)
