"""Rendering-side helpers next to the solver: the reference viewer's
draw_density (src/main.rs:146-183) computed on the device, and frame dumps for visual / regression
comparison (the reference's .gitignore mentions *.mp4 recordings; main.rs:54 prints one probe value)."""
from __future__ import annotations

import numpy as np

from ._lib import check, lib, ptr


def draw_density(c, m, y, k, z: float = 0.5) -> np.ndarray:
    """Vertex buffer of draw_density(c.density(), m.density(), y.density(), k.density(), z):
    array (w*h*4, 6) of {pos[3], color[3]} in the reference's push order."""
    w, h = c._w_, c._h_
    out = np.empty((w * h * 4, 6), np.float32)
    check(lib.fruid_draw_density(c._h, m._h, y._h, k._h, z, ptr(out)))
    return out


def density_colors(c, m, y, k) -> np.ndarray:
    """(h, w, 3) RGB colour of every cell (the `color` of main.rs:167-169)."""
    w, h = c._w_, c._h_
    out = np.empty((h, w, 3), np.float32)
    check(lib.fruid_density_colors(c._h, m._h, y._h, k._h, ptr(out)))
    return out


def draw_velocity_lines(sim, z: float = 0.5) -> np.ndarray:
    """Vertex buffer of draw_velocity_lines(sim.uv(), z) (main.rs:185-216): array (w*h*2, 6),
    {tail, tip} per cell in the reference's push order."""
    w, h = sim.width(), sim.height()
    out = np.empty((w * h * 2, 6), np.float32)
    check(lib.fruid_draw_velocity_lines(sim._h, z, ptr(out)))
    return out


def write_ppm(path: str, rgb: np.ndarray) -> None:
    """Binary PPM (P6) of an (h, w, 3) float image; colours are clamped to [0, 1] like a framebuffer."""
    img = np.nan_to_num(np.asarray(rgb, np.float32), nan=0.0, posinf=1.0, neginf=0.0)
    img8 = (np.clip(img, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
    h, w, _ = img8.shape
    with open(path, "wb") as fh:
        fh.write(b"P6\n%d %d\n255\n" % (w, h))
        fh.write(img8.tobytes())


def dump_fields(path: str, **fields) -> None:
    """Regression dump: named float32 fields into one .npz."""
    np.savez_compressed(path, **{k: np.asarray(v, np.float32) for k, v in fields.items()})
