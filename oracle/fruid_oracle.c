/*
 * fruid_oracle.c -- CPU restatement of the stable-fluids time step of
 * Masterchef365/fruid (reference: src/lib.rs:5-208).
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path (fruid_b200/) never links, imports or calls
 * anything in oracle/.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 * for this path and cannot be compiled in this
 * image (no rustc/cargo, un-vendored dependencies, no network).  The oracle is
 * therefore pinned only by (1) a line-by-line audit against src/lib.rs (every
 * function below cites the lines it restates), (2) analytic known answers
 * derived from the source text and (3) two independent restatements
 * (tests/ref_numpy.py, vectorised; tests/ref_literal.py, a statement-by-
 * statement scalar transliteration) that must agree bit-for-bit.
 *
 * Numeric contract (what Rust does, what this file must do):
 *   - f32 everywhere; only + - * / compare negate and float->usize casts.
 *   - Rust never contracts a*b+c into an FMA and never reassociates:
 *     build with -ffp-contract=off and without -ffast-math (see Makefile).
 *   - `x as usize` from f32 truncates toward zero, saturates, NaN -> 0.
 *   - Clamps are `<` / `>` comparisons (NaN passes through), not fmin/fmax.
 *   - Array2D<f32> (idek_basics @ 6a3e9c74, not vendored) is a flat Vec<f32>
 *     with index (x, y) -> x + y*width; that function lives in IDX() only.
 *
 * Two traversal orders are provided for every loop nest:
 *   order 0 = FAITHFUL: x outer / y inner exactly like src/lib.rs:53-54,
 *             90-91, 134-135, 149-150 (cache-hostile; this is what the CPU
 *             baseline times, single thread, like the reference);
 *   order 1 = FAST: y outer / x inner (+ OpenMP when compiled with -fopenmp).
 * Every loop in the reference is out-of-place and per-cell independent, so
 * both orders give bit-identical results (tests/test_oracle.py checks it).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define IDX(x, y, w) ((size_t)(x) + (size_t)(y) * (size_t)(w))

enum { B_POSITIVE = 0, B_NEGX = 1, B_NEGY = 2 }; /* src/lib.rs:19-25 */
enum { SIDE_X = 0, SIDE_Y = 1 };                 /* src/lib.rs:128-130 */

/* Rust `f as usize` (saturating, NaN -> 0), used at src/lib.rs:102,112. */
static inline size_t f32_as_usize(float f) {
    if (!(f == f)) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)f;
}

/* src/lib.rs:5-10 -- every flat element, border included. */
void fruid_oracle_add_source(float *x, const float *s, size_t w, size_t h, float dt) {
    size_t n = w * h;
    for (size_t k = 0; k < n; ++k) x[k] = x[k] + s[k] * dt;
}

/* src/lib.rs:27-44.  Note the first loop's variable runs over y (1..=ny). */
void fruid_oracle_set_bnd(int b, float *x, size_t w, size_t h) {
    size_t nx = w - 2, ny = h - 2; /* src/lib.rs:12-17 */
    for (size_t i = 1; i <= ny; ++i) {
        x[IDX(0, i, w)] = (b == B_NEGX) ? -x[IDX(1, i, w)] : x[IDX(1, i, w)];
        x[IDX(nx + 1, i, w)] = (b == B_NEGX) ? -x[IDX(nx, i, w)] : x[IDX(nx, i, w)];
    }
    for (size_t i = 1; i <= nx; ++i) {
        x[IDX(i, 0, w)] = (b == B_NEGY) ? -x[IDX(i, 1, w)] : x[IDX(i, 1, w)];
        x[IDX(i, ny + 1, w)] = (b == B_NEGY) ? -x[IDX(i, ny, w)] : x[IDX(i, ny, w)];
    }
    x[IDX(0, 0, w)] = 0.5f * (x[IDX(1, 0, w)] + x[IDX(0, 1, w)]);
    x[IDX(0, ny + 1, w)] = 0.5f * (x[IDX(1, ny + 1, w)] + x[IDX(0, ny, w)]);
    x[IDX(nx + 1, 0, w)] = 0.5f * (x[IDX(nx, 0, w)] + x[IDX(nx + 1, 1, w)]);
    x[IDX(nx + 1, ny + 1, w)] = 0.5f * (x[IDX(nx, ny + 1, w)] + x[IDX(nx + 1, ny, w)]);
}

/* One interior cell of src/lib.rs:56-65. */
static inline void lin_cell(float *out, const float *x, const float *x0, size_t i, size_t j,
                            size_t w, float a, float c) {
    float left = x[IDX(i - 1, j, w)];
    float right = x[IDX(i + 1, j, w)];
    float up = x[IDX(i, j - 1, w)];
    float down = x[IDX(i, j + 1, w)];
    float sum = left + right + up + down; /* ((l + r) + u) + d */
    float center_prev = x0[IDX(i, j, w)];
    out[IDX(i, j, w)] = (center_prev + a * sum) / c;
}

/*
 * src/lib.rs:46-71 with the hard-coded 20 (lib.rs:52) exposed as `iters`.
 * `x` and `scratch` are re-bound references that swap every sweep (lib.rs:68),
 * so after an even number of sweeps the caller's `x` holds the result; after
 * an odd number the caller's `scratch` holds it (not reference behaviour --
 * the reference only ever runs 20).
 */
void fruid_oracle_lin_solve(int b, float *x, const float *x0, float *scratch, size_t w, size_t h,
                            float a, float c, int iters, int order) {
    size_t nx = w - 2, ny = h - 2;
    for (int it = 0; it < iters; ++it) {
        if (order == 0) {
            for (size_t i = 1; i <= nx; ++i)
                for (size_t j = 1; j <= ny; ++j) lin_cell(scratch, x, x0, i, j, w, a, c);
        } else {
#pragma omp parallel for schedule(static)
            for (size_t j = 1; j <= ny; ++j)
                for (size_t i = 1; i <= nx; ++i) lin_cell(scratch, x, x0, i, j, w, a, c);
        }
        float *t = scratch; scratch = x; x = t; /* std::mem::swap, lib.rs:68 */
        fruid_oracle_set_bnd(b, x, w, h);       /* lib.rs:69 */
    }
}

/* src/lib.rs:73-78.  `dt * diff * nx as f32 * ny as f32` is left-associated. */
void fruid_oracle_diffuse(int b, float *x, const float *x0, float *scratch, size_t w, size_t h,
                          float diff, float dt, int iters, int order) {
    size_t nx = w - 2, ny = h - 2;
    float a = dt * diff * (float)nx * (float)ny;
    fruid_oracle_lin_solve(b, x, x0, scratch, w, h, a, 1.0f + 4.0f * a, iters, order);
}

/* src/lib.rs:80-82 */
static inline float mix(float a, float b, float t) { return (1.0f - t) * a + t * b; }

/* One interior cell of src/lib.rs:92-122. */
static inline void advect_cell(float *d, const float *d0, const float *u, const float *v, size_t i,
                               size_t j, size_t w, size_t nx, size_t ny, float dt0, float dt1) {
    float x = (float)i - dt0 * u[IDX(i, j, w)];
    float y = (float)j - dt1 * v[IDX(i, j, w)];
    if (x < 0.5f) x = 0.5f;
    if (x > (float)nx + 0.5f) x = (float)nx + 0.5f;
    size_t i0 = f32_as_usize(x);
    size_t i1 = i0 + 1;
    if (y < 0.5f) y = 0.5f;
    if (y > (float)ny + 0.5f) y = (float)ny + 0.5f;
    size_t j0 = f32_as_usize(y);
    size_t j1 = j0 + 1;
    float s1 = x - (float)i0;
    float t1 = y - (float)j0;
    d[IDX(i, j, w)] = mix(mix(d0[IDX(i0, j0, w)], d0[IDX(i0, j1, w)], t1),
                          mix(d0[IDX(i1, j0, w)], d0[IDX(i1, j1, w)], t1), s1);
}

/* src/lib.rs:84-126.  d must not alias d0/u/v (Rust's &mut guarantees it). */
void fruid_oracle_advect(int b, float *d, const float *d0, const float *u, const float *v, size_t w,
                         size_t h, float dt, int order) {
    size_t nx = w - 2, ny = h - 2;
    float dt0 = dt * (float)nx;
    float dt1 = dt * (float)ny;
    if (order == 0) {
        for (size_t i = 1; i <= nx; ++i)
            for (size_t j = 1; j <= ny; ++j) advect_cell(d, d0, u, v, i, j, w, nx, ny, dt0, dt1);
    } else {
#pragma omp parallel for schedule(static)
        for (size_t j = 1; j <= ny; ++j)
            for (size_t i = 1; i <= nx; ++i) advect_cell(d, d0, u, v, i, j, w, nx, ny, dt0, dt1);
    }
    fruid_oracle_set_bnd(b, d, w, h);
}

/* src/lib.rs:132-144 */
void fruid_oracle_diff_acc(float *r, const float *d, size_t w, size_t h, float scale, int side,
                           int order) {
    size_t nx = w - 2, ny = h - 2;
#define DIFF_CELL(i, j)                                                                  \
    do {                                                                                 \
        float diff = (side == SIDE_X) ? d[IDX((i) + 1, (j), w)] - d[IDX((i)-1, (j), w)]  \
                                      : d[IDX((i), (j) + 1, w)] - d[IDX((i), (j)-1, w)]; \
        r[IDX((i), (j), w)] = r[IDX((i), (j), w)] + scale * diff;                        \
    } while (0)
    if (order == 0) {
        for (size_t i = 1; i <= nx; ++i)
            for (size_t j = 1; j <= ny; ++j) DIFF_CELL(i, j);
    } else {
#pragma omp parallel for schedule(static)
        for (size_t j = 1; j <= ny; ++j)
            for (size_t i = 1; i <= nx; ++i) DIFF_CELL(i, j);
    }
#undef DIFF_CELL
}

/* src/lib.rs:146-169.  `-0.5 / nx as f32` is (-0.5)/(nx as f32). */
void fruid_oracle_project(float *u, float *v, float *p, float *div, float *scratch, size_t w,
                          size_t h, int iters, int order) {
    size_t nx = w - 2, ny = h - 2;
    if (order == 0) {
        for (size_t i = 1; i <= nx; ++i)
            for (size_t j = 1; j <= ny; ++j) {
                div[IDX(i, j, w)] = 0.0f;
                p[IDX(i, j, w)] = 0.0f;
            }
    } else {
        for (size_t j = 1; j <= ny; ++j)
            for (size_t i = 1; i <= nx; ++i) {
                div[IDX(i, j, w)] = 0.0f;
                p[IDX(i, j, w)] = 0.0f;
            }
    }
    fruid_oracle_diff_acc(div, u, w, h, -0.5f / (float)nx, SIDE_X, order);
    fruid_oracle_diff_acc(div, v, w, h, -0.5f / (float)ny, SIDE_Y, order);

    fruid_oracle_set_bnd(B_POSITIVE, div, w, h);
    fruid_oracle_set_bnd(B_POSITIVE, p, w, h);

    fruid_oracle_lin_solve(B_POSITIVE, p, div, scratch, w, h, 1.0f, 4.0f, iters, order);

    fruid_oracle_diff_acc(u, p, w, h, -0.5f * (float)nx, SIDE_X, order);
    fruid_oracle_diff_acc(v, p, w, h, -0.5f * (float)ny, SIDE_Y, order);

    fruid_oracle_set_bnd(B_NEGX, u, w, h);
    fruid_oracle_set_bnd(B_NEGY, v, w, h);
}

/* src/lib.rs:171-179 (the SWAPs are commented out in the reference). */
void fruid_oracle_dens_step(float *x, float *x0, const float *u, const float *v, float *scratch,
                            size_t w, size_t h, float diff, float dt, int iters, int order) {
    fruid_oracle_add_source(x, x0, w, h, dt);
    fruid_oracle_diffuse(B_POSITIVE, x0, x, scratch, w, h, diff, dt, iters, order);
    fruid_oracle_advect(B_POSITIVE, x, x0, u, v, w, h, dt, order);
}

/* src/lib.rs:181-208.  On exit u0 holds the pressure and v0 the divergence of
 * the second projection (lib.rs:207 -> lib.rs:146). */
void fruid_oracle_vel_step(float *u, float *v, float *u0, float *v0, float *scratch, size_t w,
                           size_t h, float visc, float dt, int iters, int order) {
    fruid_oracle_add_source(u, u0, w, h, dt);
    fruid_oracle_add_source(v, v0, w, h, dt);
    fruid_oracle_diffuse(B_NEGX, u0, u, scratch, w, h, visc, dt, iters, order);
    fruid_oracle_diffuse(B_NEGY, v0, v, scratch, w, h, visc, dt, iters, order);
    fruid_oracle_project(u0, v0, u, v, scratch, w, h, iters, order);
    fruid_oracle_advect(B_NEGX, u, u0, u0, v0, w, h, dt, order);
    fruid_oracle_advect(B_NEGY, v, v0, u0, v0, w, h, dt, order);
    fruid_oracle_project(u, v, u0, v0, scratch, w, h, iters, order);
}

int fruid_oracle_has_openmp(void) {
#ifdef _OPENMP
    return 1;
#else
    return 0;
#endif
}

/* ---- "next" row f-2: draw_density, src/main.rs:146-183 -------------------------
 * Per cell (i outer, j inner): CMY dye -> RGB colour log2((1 - f - k)/max(|cmy|, 1) + 2) and four
 * vertices {pos[3], color[3]} (tl, tr, bl, br).  `verts` holds w*h*4*6 floats in push order.
 * log2 is libm's log2f here and CUDA's log2f on the device: colours agree to a couple of ulps,
 * positions bit for bit. */
#include <math.h>
void fruid_oracle_draw_density(const float *c, const float *m, const float *y, const float *k, size_t w,
                               size_t h, float z, float *verts) {
    float cell_width = 2.0f / (float)w;
    float cell_height = 2.0f / (float)h;
    size_t n = 0;
    for (size_t i = 0; i < w; ++i) {
        float i_frac = ((float)i / (float)w) * 2.0f - 1.0f;
        for (size_t j = 0; j < h; ++j) {
            float j_frac = ((float)j / (float)h) * 2.0f - 1.0f;
            float kk = k[IDX(i, j, w)];
            float cmy[3] = {c[IDX(i, j, w)], m[IDX(i, j, w)], y[IDX(i, j, w)]};
            float sum = 0.0f;
            for (int d = 0; d < 3; ++d) sum = sum + cmy[d] * cmy[d];
            float total = sqrtf(sum);
            float denom = (total > 1.0f || total != total) ? total : 1.0f; /* f32::max(total, 1.) */
            if (total != total) denom = 1.0f;                              /* NaN.max(1.) == 1. */
            float col[3];
            for (int d = 0; d < 3; ++d) col[d] = log2f((1.0f - cmy[d] - kk) / denom + 2.0f);
            float dx[4] = {0.0f, cell_width, 0.0f, cell_width};
            float dy[4] = {0.0f, 0.0f, cell_height, cell_height};
            for (int q = 0; q < 4; ++q) {
                verts[n++] = i_frac + dx[q];
                verts[n++] = j_frac + dy[q];
                verts[n++] = z;
                verts[n++] = col[0];
                verts[n++] = col[1];
                verts[n++] = col[2];
            }
        }
    }
}

/* ---- "next" row f-4: draw_velocity_lines, src/main.rs:185-216 (init-only debug
 * visualisation, main.rs:56).  Per cell (i outer, j inner) two vertices {pos[3], color[3]}: the
 * tail at the cell centre and the tip at tail + (u, v) * (2*cell_height / speed), both coloured
 * [speed; 3].  `verts` holds w*h*2*6 floats in push order.  `x.powf(2.)` is restated as x*x (LLVM
 * folds pow(x, 2.0) into a multiply; the values agree for every input either way up to libm's
 * rounding of powf).  speed == 0 gives len = inf and a NaN tip (0*inf), as in the reference. */
void fruid_oracle_draw_velocity_lines(const float *u, const float *v, size_t w, size_t h, float z, float *verts) {
    float cell_width = 2.0f / (float)w;
    float cell_height = 2.0f / (float)h;
    size_t n = 0;
    for (size_t i = 0; i < w; ++i) {
        float i_frac = ((float)i / (float)w) * 2.0f - 1.0f;
        for (size_t j = 0; j < h; ++j) {
            float j_frac = ((float)j / (float)h) * 2.0f - 1.0f;
            float uu = u[IDX(i, j, w)];
            float vv = v[IDX(i, j, w)];
            float speed = sqrtf(uu * uu + vv * vv);
            float tail_x = i_frac + cell_width / 2.0f;
            float tail_y = j_frac + cell_height / 2.0f;
            float len = cell_height * 2.0f / speed;
            float pos[2][2] = {{tail_x, tail_y}, {tail_x + uu * len, tail_y + vv * len}};
            for (int q = 0; q < 2; ++q) {
                verts[n++] = pos[q][0];
                verts[n++] = pos[q][1];
                verts[n++] = z;
                verts[n++] = speed;
                verts[n++] = speed;
                verts[n++] = speed;
            }
        }
    }
}
