"""CPU-only: pins the oracle (oracle/fruid_oracle.c).  The reference has no tests or golden
vectors for this path, so the pins are: an independent numpy restatement
(tests/ref_numpy.py), analytic known answers derived from src/lib.rs, traversal-order
independence, and the committed golden fixtures (tests/golden/, made by make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

import ref_numpy as R
from conftest import assert_bits, rand_field
from fruid_b200_scene_import import Scene
from oracle_api import ODensitySim, OFluidSim

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SIZES = [(3, 3), (3, 9), (9, 3), (4, 7), (34, 27), (65, 130)]


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("visc", [0.0, 3e-4])
def test_vel_step_matches_numpy_restatement(oracle_mod, w, h, visc):
    rng = np.random.default_rng(w * 1000 + h)
    A = [rand_field(rng, w, h) for _ in range(5)]
    B = [a.copy() for a in A]
    for _ in range(2):
        oracle_mod.lib().fruid_oracle_vel_step(*A, w, h, visc, 0.1, 20, oracle_mod.FAITHFUL)
        R.vel_step(*B, visc, 0.1, 20)
    for a, b, n in zip(A[:4], B[:4], ("u", "v", "u_prev", "v_prev")):
        assert_bits(a, b, n)


@pytest.mark.parametrize("w,h", SIZES)
def test_dens_step_matches_numpy_restatement(oracle_mod, w, h):
    rng = np.random.default_rng(7 + w)
    x, x0, s = (rand_field(rng, w, h) for _ in range(3))
    u, v = rand_field(rng, w, h, 3.0), rand_field(rng, w, h, 3.0)
    x2, x02, s2 = x.copy(), x0.copy(), s.copy()
    oracle_mod.lib().fruid_oracle_dens_step(x, x0, u, v, s, w, h, 1e-4, 0.1, 20, oracle_mod.FAITHFUL)
    R.dens_step(x2, x02, u, v, s2, 1e-4, 0.1, 20)
    assert_bits(x, x2, "dens")
    assert_bits(x0, x02, "dens_prev")


@pytest.mark.parametrize("w,h", [(5, 4), (40, 33)])
def test_traversal_orders_and_openmp_agree(oracle_mod, w, h):
    rng = np.random.default_rng(3)
    A = [rand_field(rng, w, h) for _ in range(5)]
    B = [a.copy() for a in A]
    C = [a.copy() for a in A]
    oracle_mod.lib().fruid_oracle_vel_step(*A, w, h, 1e-4, 0.05, 20, oracle_mod.FAITHFUL)
    oracle_mod.lib().fruid_oracle_vel_step(*B, w, h, 1e-4, 0.05, 20, oracle_mod.FAST)
    oracle_mod.lib(True).fruid_oracle_vel_step(*C, w, h, 1e-4, 0.05, 20, oracle_mod.FAST)
    for a, b, c in zip(A, B, C):
        assert_bits(a, b)
        assert_bits(a, c)


# ---- analytic known answers derived from the source text ----------------
def test_zero_fields_stay_zero(oracle_mod):
    f = oracle_mod.Fluid(17, 12)
    f.step(0.1, 1e-3)
    for a in (f.u, f.v, f.u_prev, f.v_prev):
        assert not a.any()


def test_set_bnd_known_answers(oracle_mod):
    w, h = 9, 7
    rng = np.random.default_rng(0)
    for b in (oracle_mod.POSITIVE, oracle_mod.NEGX, oracle_mod.NEGY):
        x = rand_field(rng, w, h)
        oracle_mod.lib().fruid_oracle_set_bnd(b, x, w, h)
        sx = -1 if b == oracle_mod.NEGX else 1
        sy = -1 if b == oracle_mod.NEGY else 1
        assert np.array_equal(x[1:-1, 0], sx * x[1:-1, 1])          # lib.rs:31
        assert np.array_equal(x[1:-1, -1], sx * x[1:-1, -2])        # lib.rs:32
        assert np.array_equal(x[0, 1:-1], sy * x[1, 1:-1])          # lib.rs:36
        assert np.array_equal(x[-1, 1:-1], sy * x[-2, 1:-1])        # lib.rs:37
        if b == oracle_mod.POSITIVE:  # corner = adjacent interior value
            assert x[0, 0] == x[1, 1] and x[-1, -1] == x[-2, -2]
        else:                         # one edge mirrored, the other negated: 0.5*(a + -a) = +0
            for c in (x[0, 0], x[0, -1], x[-1, 0], x[-1, -1]):
                assert c == 0.0 and not np.signbit(c)


def test_advect_with_zero_velocity_is_identity_on_interior(oracle_mod):
    w, h = 20, 15
    rng = np.random.default_rng(5)
    d0 = rand_field(rng, w, h)
    d = np.zeros_like(d0)
    z = np.zeros_like(d0)
    oracle_mod.lib().fruid_oracle_advect(0, d, d0, z, z, w, h, 0.1, 0)
    assert np.array_equal(d[1:-1, 1:-1], d0[1:-1, 1:-1])  # s1 = t1 = 0 -> 1*a + 0*b


def test_inviscid_diffuse_returns_x0(oracle_mod):
    w, h = 12, 10
    rng = np.random.default_rng(6)
    x, x0, s = rand_field(rng, w, h), rand_field(rng, w, h), rand_field(rng, w, h)
    oracle_mod.lib().fruid_oracle_diffuse(0, x, x0, s, w, h, 0.0, 0.1, 20, 0)
    assert np.array_equal(x[1:-1, 1:-1], x0[1:-1, 1:-1])  # a = 0, c = 1


def test_project_leaves_uniform_flow_interior_unchanged(oracle_mod):
    w, h = 16, 16
    u = np.full((h, w), 0.25, np.float32)
    v = np.full((h, w), -0.5, np.float32)
    p, div, s = (np.ones((h, w), np.float32) for _ in range(3))
    u0, v0 = u.copy(), v.copy()
    oracle_mod.lib().fruid_oracle_project(u, v, p, div, s, w, h, 20, 0)
    assert not div[1:-1, 1:-1].any()          # divergence-free
    assert not p.any()                         # stale p is dead: zeroed at lib.rs:152
    assert np.array_equal(u[1:-1, 1:-1], u0[1:-1, 1:-1])
    assert np.array_equal(v[1:-1, 1:-1], v0[1:-1, 1:-1])


def test_step_leaves_pressure_and_divergence_in_prev_fields(oracle_mod):
    """After FluidSim::step, u_prev = p and v_prev = div of the 2nd projection (lib.rs:207)."""
    w, h = 14, 11
    rng = np.random.default_rng(8)
    f = oracle_mod.Fluid(w, h, order=oracle_mod.FAITHFUL)
    for a in (f.u, f.v, f.u_prev, f.v_prev):
        a[:] = rand_field(rng, w, h)
    u, v, u0, v0, s = (a.copy() for a in (f.u, f.v, f.u_prev, f.v_prev, f.scratch))
    f.step(0.1, 0.0)
    L = oracle_mod.lib()
    L.fruid_oracle_add_source(u, u0, w, h, 0.1)
    L.fruid_oracle_add_source(v, v0, w, h, 0.1)
    L.fruid_oracle_diffuse(1, u0, u, s, w, h, 0.0, 0.1, 20, 0)
    L.fruid_oracle_diffuse(2, v0, v, s, w, h, 0.0, 0.1, 20, 0)
    L.fruid_oracle_project(u0, v0, u, v, s, w, h, 20, 0)
    L.fruid_oracle_advect(1, u, u0, u0, v0, w, h, 0.1, 0)
    L.fruid_oracle_advect(2, v, v0, u0, v0, w, h, 0.1, 0)
    p, div = u0, v0
    L.fruid_oracle_project(u, v, p, div, s, w, h, 20, 0)
    assert_bits(f.u_prev, p, "u_prev == pressure")
    assert_bits(f.v_prev, div, "v_prev == divergence")
    assert_bits(f.u, u, "u")


def test_nan_and_inf_follow_rust_cast_semantics(oracle_mod):
    """NaN velocity passes the clamps (lib.rs:95-110) and `NaN as usize` = 0."""
    w, h = 8, 8
    d0 = np.arange(w * h, dtype=np.float32).reshape(h, w)
    u = np.zeros((h, w), np.float32)
    v = np.zeros((h, w), np.float32)
    u[3, 3] = np.nan
    v[4, 4] = np.inf
    u[5, 5] = -np.inf
    d = np.zeros_like(d0)
    oracle_mod.lib().fruid_oracle_advect(0, d, d0, u, v, w, h, 0.1, 0)
    d2 = np.zeros_like(d0)
    R.advect(0, d2, d0, u, v, 0.1)
    assert_bits(d, d2)
    assert np.isnan(d[3, 3])
    assert d[4, 4] == d0[1, 4] * 0.5 + d0[0, 4] * 0.5  # y clamps to 0.5 -> j0=0,t1=.5


# ---- golden fixtures -------------------------------------------------------------------
@pytest.mark.parametrize("name", ["steps_37x29_visc.npz", "steps_30x41_inviscid.npz"])
def test_golden_steps(oracle_mod, name):
    g = np.load(os.path.join(GOLD, name))
    w, h = int(g["w"]), int(g["h"])
    f = oracle_mod.Fluid(w, h)
    d = oracle_mod.Density(w, h)
    for k, a in (("u", f.u), ("v", f.v), ("u_prev", f.u_prev), ("v_prev", f.v_prev), ("dens", d.dens),
                 ("dens_prev", d.dens_prev)):
        a[:] = g["in_" + k]
    for _ in range(int(g["nsteps"])):
        f.step(float(g["dt"]), float(g["visc"]))
        d.step((f.u, f.v), float(g["dt"]), float(g["diff"]))
    for k, a in (("u", f.u), ("v", f.v), ("u_prev", f.u_prev), ("v_prev", f.v_prev), ("dens", d.dens),
                 ("dens_prev", d.dens_prev)):
        assert_bits(a, g["out_" + k], k)


def _scene_fields(s):
    u, v = s.sim.uv()
    p, dv = s.sim.uv_mut()
    out = {"u": u.numpy(), "v": v.numpy(), "u_prev": p.numpy(), "v_prev": dv.numpy()}
    for i, d in enumerate(s.dyes):
        out[f"dens{i}"] = d.density().numpy()
        out[f"dens_prev{i}"] = d.density_mut().numpy()
    return out


def test_golden_scene_small(oracle_mod):
    g = np.load(os.path.join(GOLD, "scene_64x48_f6.npz"))
    s = Scene(OFluidSim, ODensitySim, 64, 48)
    s.run(6)
    for k, a in _scene_fields(s).items():
        assert_bits(a, g[k], k)


def test_golden_scene_checksums_128(oracle_mod):
    sums = json.load(open(os.path.join(GOLD, "scene_checksums.json")))["128x128_f100"]
    s = Scene(OFluidSim, ODensitySim, 128, 128)
    s.run(100)
    for k, a in _scene_fields(s).items():
        assert hashlib.sha256(a.tobytes()).hexdigest() == sums["sha256"][k], k


def test_draw_density_known_answers(oracle_mod):
    """main.rs:146-183: with no dye every colour channel is log2((1 - 0 - 0)/max(0, 1) + 2) = log2(3); cell
    (i, j) is the (i*h + j)-th quad, its corners at i_frac, i_frac + 2/w; z is passed through."""
    w, h = 6, 4
    z = np.zeros((h, w), np.float32)
    v = np.zeros(w * h * 24, np.float32)
    oracle_mod.lib().fruid_oracle_draw_density(z, z, z, z, w, h, 0.25, v)
    v = v.reshape(w, h, 4, 6)
    assert np.allclose(v[..., 3:], np.log2(3.0), rtol=1e-6)
    assert np.all(v[..., 2] == 0.25)
    i, j = 2, 3
    x0 = np.float32(np.float32(i) / np.float32(w)) * np.float32(2) - np.float32(1)
    y0 = np.float32(np.float32(j) / np.float32(h)) * np.float32(2) - np.float32(1)
    assert v[i, j, 0, 0] == x0 and v[i, j, 0, 1] == y0                     # tl
    assert v[i, j, 1, 0] == x0 + np.float32(2.0 / w) and v[i, j, 1, 1] == y0  # tr
    assert v[i, j, 2, 0] == x0 and v[i, j, 2, 1] == y0 + np.float32(2.0 / h)  # bl
    # heavy dye: colours divide by |cmy|
    c = np.full((h, w), 3.0, np.float32)
    oracle_mod.lib().fruid_oracle_draw_density(c, z, z, z, w, h, 0.0, v.reshape(-1))
    assert np.allclose(v[..., 3], np.log2((1 - 3.0) / 3.0 + 2), rtol=1e-6)
    assert np.allclose(v[..., 4], np.log2(1 / 3.0 + 2), rtol=1e-6)


def test_draw_velocity_lines_known_answers(oracle_mod):
    """main.rs:185-216: the tail sits at the cell centre, the tip at tail + (u, v) * 2*cell_height/speed
    (so every line is 2 cell heights long), colour = speed; zero velocity gives a NaN tip (0 * inf)."""
    w, h = 8, 4
    u = np.zeros((h, w), np.float32)
    v = np.zeros((h, w), np.float32)
    u[1, 2], v[1, 2] = 3.0, 4.0    # cell (i=2, j=1): speed 5
    u[3, 5] = -2.0                 # cell (i=5, j=3): speed 2, pointing to -x
    out = np.zeros(w * h * 12, np.float32)
    oracle_mod.lib().fruid_oracle_draw_velocity_lines(u, v, w, h, 0.75, out)
    out = out.reshape(w, h, 2, 6)
    assert np.all(out[..., 2] == 0.75)
    tail, tip = out[2, 1, 0], out[2, 1, 1]
    assert tail[0] == np.float32(2 / 8 * 2 - 1 + 1 / 8) and tail[1] == np.float32(1 / 4 * 2 - 1 + 1 / 4)
    assert np.all(tail[3:] == 5.0) and np.all(tip[3:] == 5.0)
    ln = np.float32(0.5) * np.float32(2) / np.float32(5)
    assert tip[0] == tail[0] + np.float32(3) * ln and tip[1] == tail[1] + np.float32(4) * ln
    tail, tip = out[5, 3, 0], out[5, 3, 1]
    assert tip[0] == tail[0] - np.float32(1.0) and tip[1] == tail[1]      # 2 * (2*0.5/2) = 1 to the left
    assert np.all(out[0, 0, :, 3:] == 0.0) and np.isnan(out[0, 0, 1, :2]).all() and not np.isnan(out[0, 0, 0]).any()


@pytest.mark.parametrize("w,h,visc", [(7, 6, 0.0), (6, 9, 2e-3), (3, 3, 1e-3)])
def test_oracle_matches_literal_transliteration(oracle_mod, w, h, visc):
    """tests/ref_literal.py is a statement-by-statement Python transliteration of src/lib.rs (same names,
    loop order, reference re-binding); the C oracle must agree with it bit for bit on whole steps."""
    import ref_literal as RL
    rng = np.random.default_rng(w * 31 + h)
    init = [rand_field(rng, w, h, 2.0) for _ in range(7)]
    A = [a.copy() for a in init[:5]]
    L = [RL.Array2D.from_numpy(a) for a in init[:5]]
    dA = [a.copy() for a in init[5:]] + [np.zeros((h, w), np.float32)]
    dL = [RL.Array2D.from_numpy(a) for a in dA]
    with np.errstate(all="ignore"):
        for _ in range(2):
            oracle_mod.lib().fruid_oracle_vel_step(*A, w, h, visc, 0.1, 20, oracle_mod.FAITHFUL)
            RL.vel_step(*L, visc, 0.1)
            oracle_mod.lib().fruid_oracle_dens_step(dA[0], dA[1], A[0], A[1], dA[2], w, h, visc, 0.1, 20, oracle_mod.FAITHFUL)
            RL.dens_step(dL[0], dL[1], L[0], L[1], dL[2], visc, 0.1)
    for a, l, n in zip(A[:4] + dA[:2], L[:4] + dL[:2], ("u", "v", "u_prev", "v_prev", "dens", "dens_prev")):
        assert_bits(a, l.to_numpy(), n)
