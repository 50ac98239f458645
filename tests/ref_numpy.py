"""Independent numpy restatement of src/lib.rs:5-208 (second opinion on the oracle).

TEST INFRASTRUCTURE.  Written separately from oracle/fruid_oracle.c, straight
from the reference text, as whole-array float32 expressions.  numpy evaluates
each float32 ufunc as one correctly rounded IEEE operation (no FMA contraction,
no reassociation), which is what Rust does per scalar; because every loop of the
reference is out-of-place and per-cell independent the vectorised form is
bit-identical to the scalar loops.  Arrays are (h, w) float32, [y, x].
"""
import numpy as np

F = np.float32
POSITIVE, NEGX, NEGY = 0, 1, 2


def add_source(x, s, dt):  # lib.rs:5-10
    x += s * F(dt)


def set_bnd(b, x):  # lib.rs:27-44 -- sequential: columns, rows, then corners
    h, w = x.shape
    nx, ny = w - 2, h - 2
    sx = F(-1.0) if b == NEGX else F(1.0)
    sy = F(-1.0) if b == NEGY else F(1.0)
    if ny >= 1:
        x[1:ny + 1, 0] = -x[1:ny + 1, 1] if b == NEGX else x[1:ny + 1, 1]
        x[1:ny + 1, nx + 1] = -x[1:ny + 1, nx] if b == NEGX else x[1:ny + 1, nx]
    if nx >= 1:
        x[0, 1:nx + 1] = -x[1, 1:nx + 1] if b == NEGY else x[1, 1:nx + 1]
        x[ny + 1, 1:nx + 1] = -x[ny, 1:nx + 1] if b == NEGY else x[ny, 1:nx + 1]
    del sx, sy
    x[0, 0] = F(0.5) * (x[0, 1] + x[1, 0])
    x[ny + 1, 0] = F(0.5) * (x[ny + 1, 1] + x[ny, 0])
    x[0, nx + 1] = F(0.5) * (x[0, nx] + x[1, nx + 1])
    x[ny + 1, nx + 1] = F(0.5) * (x[ny + 1, nx] + x[ny, nx + 1])


def lin_solve(b, x, x0, scratch, a, c, iters=20):  # lib.rs:46-71
    a, c = F(a), F(c)
    for _ in range(iters):
        left = x[1:-1, :-2]
        right = x[1:-1, 2:]
        up = x[:-2, 1:-1]
        down = x[2:, 1:-1]
        s = ((left + right) + up) + down
        scratch[1:-1, 1:-1] = (x0[1:-1, 1:-1] + a * s) / c
        x, scratch = scratch, x
        set_bnd(b, x)
    return x, scratch  # x = newest iterate


def diffuse(b, x, x0, scratch, diff, dt, iters=20):  # lib.rs:73-78
    h, w = x.shape
    a = F(F(F(F(dt) * F(diff)) * F(w - 2)) * F(h - 2))
    return lin_solve(b, x, x0, scratch, a, F(1.0) + F(4.0) * a, iters)


def mix(a, b, t):  # lib.rs:80-82
    return (F(1.0) - t) * a + t * b


def advect(b, d, d0, u, v, dt):  # lib.rs:84-126
    h, w = d.shape
    nx, ny = w - 2, h - 2
    dt0 = F(dt) * F(nx)
    dt1 = F(dt) * F(ny)
    ii = np.arange(1, nx + 1, dtype=np.float32)[None, :]
    jj = np.arange(1, ny + 1, dtype=np.float32)[:, None]
    with np.errstate(all="ignore"):
        x = ii - dt0 * u[1:-1, 1:-1]
        y = jj - dt1 * v[1:-1, 1:-1]
        x = np.where(x < F(0.5), F(0.5), x)
        x = np.where(x > F(nx) + F(0.5), F(nx) + F(0.5), x)
        y = np.where(y < F(0.5), F(0.5), y)
        y = np.where(y > F(ny) + F(0.5), F(ny) + F(0.5), y)
        i0 = np.where(np.isnan(x), 0, np.nan_to_num(x, nan=0.0)).astype(np.int64)
        j0 = np.where(np.isnan(y), 0, np.nan_to_num(y, nan=0.0)).astype(np.int64)
        i1, j1 = i0 + 1, j0 + 1
        s1 = x - i0.astype(np.float32)
        t1 = y - j0.astype(np.float32)
        d0c = d0.copy() if np.shares_memory(d, d0) else d0
        d[1:-1, 1:-1] = mix(mix(d0c[j0, i0], d0c[j1, i0], t1), mix(d0c[j0, i1], d0c[j1, i1], t1), s1)
    set_bnd(b, d)


def diff_acc(r, d, scale, side):  # lib.rs:132-144
    if side == "x":
        diff = d[1:-1, 2:] - d[1:-1, :-2]
    else:
        diff = d[2:, 1:-1] - d[:-2, 1:-1]
    r[1:-1, 1:-1] += F(scale) * diff


def project(u, v, p, div, scratch, iters=20):  # lib.rs:146-169
    h, w = u.shape
    nx, ny = w - 2, h - 2
    div[1:-1, 1:-1] = 0
    p[1:-1, 1:-1] = 0
    with np.errstate(all="ignore"):
        diff_acc(div, u, F(-0.5) / F(nx), "x")
        diff_acc(div, v, F(-0.5) / F(ny), "y")
    set_bnd(POSITIVE, div)
    set_bnd(POSITIVE, p)
    newest, other = lin_solve(POSITIVE, p, div, scratch, 1.0, 4.0, iters)
    assert newest is p or iters % 2 == 1
    diff_acc(u, p, F(-0.5) * F(nx), "x")
    diff_acc(v, p, F(-0.5) * F(ny), "y")
    set_bnd(NEGX, u)
    set_bnd(NEGY, v)


def dens_step(x, x0, u, v, scratch, diff, dt, iters=20):  # lib.rs:171-179
    add_source(x, x0, dt)
    diffuse(POSITIVE, x0, x, scratch, diff, dt, iters)
    advect(POSITIVE, x, x0, u, v, dt)


def vel_step(u, v, u0, v0, scratch, visc, dt, iters=20):  # lib.rs:181-208
    add_source(u, u0, dt)
    add_source(v, v0, dt)
    diffuse(NEGX, u0, u, scratch, visc, dt, iters)
    diffuse(NEGY, v0, v, scratch, visc, dt, iters)
    project(u0, v0, u, v, scratch, iters)
    advect(NEGX, u, u0, u0, v0, dt)
    advect(NEGY, v, v0, u0, v0, dt)
    project(u, v, u0, v0, scratch, iters)
