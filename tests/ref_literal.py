"""Line-by-line transliteration of the reference's src/lib.rs:5-208 into scalar Python (third opinion).

TEST INFRASTRUCTURE.  Deliberately literal: same function names, same argument names and order,
same loop variables and loop order, same statement order, the same reference re-binding in lin_solve,
one Python statement per Rust statement, so that the two texts can be audited side by side.  Every
arithmetic result is rounded to f32 immediately (np.float32 scalars: numpy computes f32 op f32 in
f32), which is what Rust's f32 arithmetic does; nothing is vectorised.  Slow by design -- use on tiny
grids only.
"""
import numpy as np

F = np.float32


class Array2D:
    """idek_basics::Array2D<f32>: width + flat Vec<f32>; index (x, y) -> x + y*width."""

    def __init__(self, width, height):
        self._width = width
        self._data = [F(0.0)] * (width * height)

    def width(self):
        return self._width

    def height(self):
        return len(self._data) // self._width

    def __getitem__(self, xy):
        x, y = xy
        assert 0 <= x < self._width and 0 <= y < self.height()  # Rust panics out of bounds
        return self._data[x + y * self._width]

    def __setitem__(self, xy, v):
        x, y = xy
        assert 0 <= x < self._width and 0 <= y < self.height()
        self._data[x + y * self._width] = F(v)

    def to_numpy(self):
        return np.array(self._data, np.float32).reshape(self.height(), self._width)

    @classmethod
    def from_numpy(cls, a):
        o = cls(a.shape[1], a.shape[0])
        o._data = [F(v) for v in a.reshape(-1)]
        return o


POSITIVE, NEG_X, NEG_Y = 0, 1, 2  # enum Bounds, lib.rs:19-25


def as_usize(f):  # Rust `f as usize`: truncate toward zero, saturate, NaN -> 0
    if f != f or f <= 0:
        return 0
    return int(f) if f < 1.8446744e19 else 2 ** 64 - 1


def add_source(x, s, dt):  # lib.rs:5-10
    for k in range(len(x._data)):
        x._data[k] = F(x._data[k] + F(s._data[k] * F(dt)))


def inner_size(x):  # lib.rs:12-17
    return (x.width() - 2, x.height() - 2)


def set_bnd(b, x):  # lib.rs:27-44
    nx, ny = inner_size(x)
    for i in range(1, ny + 1):
        x[(0, i)] = -x[(1, i)] if b == NEG_X else x[(1, i)]
        x[(nx + 1, i)] = -x[(nx, i)] if b == NEG_X else x[(nx, i)]
    for i in range(1, nx + 1):
        x[(i, 0)] = -x[(i, 1)] if b == NEG_Y else x[(i, 1)]
        x[(i, ny + 1)] = -x[(i, ny)] if b == NEG_Y else x[(i, ny)]
    x[(0, 0)] = F(0.5) * F(x[(1, 0)] + x[(0, 1)])
    x[(0, ny + 1)] = F(0.5) * F(x[(1, ny + 1)] + x[(0, ny)])
    x[(nx + 1, 0)] = F(0.5) * F(x[(nx, 0)] + x[(nx + 1, 1)])
    x[(nx + 1, ny + 1)] = F(0.5) * F(x[(nx, ny + 1)] + x[(nx + 1, ny)])


def lin_solve(b, x, x0, scratch, a, c):  # lib.rs:46-71
    nx, ny = inner_size(x)
    a, c = F(a), F(c)
    for _ in range(20):
        for i in range(1, nx + 1):
            for j in range(1, ny + 1):
                left = x[(i - 1, j)]
                right = x[(i + 1, j)]
                up = x[(i, j - 1)]
                down = x[(i, j + 1)]
                sum_ = F(F(F(left + right) + up) + down)
                center_prev = x0[(i, j)]
                scratch[(i, j)] = F(F(center_prev + F(a * sum_)) / c)
        scratch, x = x, scratch  # std::mem::swap(&mut scratch, &mut x): swaps the local references
        set_bnd(b, x)


def diffuse(b, x, x0, scratch, diff, dt):  # lib.rs:73-78
    nx, ny = inner_size(x)
    a = F(F(F(F(dt) * F(diff)) * F(nx)) * F(ny))
    lin_solve(b, x, x0, scratch, a, F(F(1.0) + F(F(4.0) * a)))


def mix(a, b, t):  # lib.rs:80-82
    return F(F(F(F(1.0) - t) * a) + F(t * b))


def advect(b, d, d0, u, v, dt):  # lib.rs:84-126
    nx, ny = inner_size(d)
    dt0 = F(F(dt) * F(nx))
    dt1 = F(F(dt) * F(ny))
    with np.errstate(all="ignore"):
        for i in range(1, nx + 1):
            for j in range(1, ny + 1):
                x = F(F(i) - F(dt0 * u[(i, j)]))
                y = F(F(j) - F(dt1 * v[(i, j)]))
                if x < F(0.5):
                    x = F(0.5)
                if x > F(F(nx) + F(0.5)):
                    x = F(F(nx) + F(0.5))
                i0 = as_usize(x)
                i1 = i0 + 1
                if y < F(0.5):
                    y = F(0.5)
                if y > F(F(ny) + F(0.5)):
                    y = F(F(ny) + F(0.5))
                j0 = as_usize(y)
                j1 = j0 + 1
                s1 = F(x - F(i0))
                t1 = F(y - F(j0))
                d[(i, j)] = mix(mix(d0[(i0, j0)], d0[(i0, j1)], t1), mix(d0[(i1, j0)], d0[(i1, j1)], t1), s1)
    set_bnd(b, d)


def diff_acc(r, d, scale, side):  # lib.rs:132-144
    nx, ny = inner_size(d)
    for i in range(1, nx + 1):
        for j in range(1, ny + 1):
            if side == "X":
                diff = F(d[(i + 1, j)] - d[(i - 1, j)])
            else:
                diff = F(d[(i, j + 1)] - d[(i, j - 1)])
            r[(i, j)] = F(r[(i, j)] + F(F(scale) * diff))


def project(u, v, p, div, scratch):  # lib.rs:146-169
    nx, ny = inner_size(u)
    for i in range(1, nx + 1):
        for j in range(1, ny + 1):
            div[(i, j)] = F(0.0)
            p[(i, j)] = F(0.0)
    diff_acc(div, u, F(F(-0.5) / F(nx)), "X")
    diff_acc(div, v, F(F(-0.5) / F(ny)), "Y")
    set_bnd(POSITIVE, div)
    set_bnd(POSITIVE, p)
    lin_solve(POSITIVE, p, div, scratch, 1.0, 4.0)
    diff_acc(u, p, F(F(-0.5) * F(nx)), "X")
    diff_acc(v, p, F(F(-0.5) * F(ny)), "Y")
    set_bnd(NEG_X, u)
    set_bnd(NEG_Y, v)


def dens_step(x, x0, u, v, scratch, diff, dt):  # lib.rs:171-179
    add_source(x, x0, dt)
    diffuse(POSITIVE, x0, x, scratch, diff, dt)
    advect(POSITIVE, x, x0, u, v, dt)


def vel_step(u, v, u0, v0, scratch, visc, dt):  # lib.rs:181-208
    add_source(u, u0, dt)
    add_source(v, v0, dt)
    diffuse(NEG_X, u0, u, scratch, visc, dt)
    diffuse(NEG_Y, v0, v, scratch, visc, dt)
    project(u0, v0, u, v, scratch)
    advect(NEG_X, u, u0, u0, v0, dt)
    advect(NEG_Y, v, v0, u0, v0, dt)
    project(u, v, u0, v0, scratch)
