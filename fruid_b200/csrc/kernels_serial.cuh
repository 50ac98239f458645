// kernels_serial.cuh -- single-thread kernels for DEGENERATE grids only (w == 2 or
// h == 2, i.e. an empty interior, src/lib.rs:12-17).  With nx == 0 or ny == 0 the border
// writes of set_bnd (src/lib.rs:27-44) read cells that set_bnd itself has just written,
// so the result depends on the statement order of the reference; one thread replays
// that order literally.  Every interior loop nest of the reference is empty for these
// sizes, so only set_bnd (and the all-cells add_source) has any effect.
#pragma once
#include "common.cuh"

namespace fruid {

__device__ inline void serial_set_bnd(float* x, const Geom& g, int b) {
    const int nx = g.nx, ny = g.ny;
    for (int i = 1; i <= ny; ++i) {  // lib.rs:30-33
        x[at(g, 0, i)] = sgn(b == B_NEGX, x[at(g, 1, i)]);
        x[at(g, nx + 1, i)] = sgn(b == B_NEGX, x[at(g, nx, i)]);
    }
    for (int i = 1; i <= nx; ++i) {  // lib.rs:35-38
        x[at(g, i, 0)] = sgn(b == B_NEGY, x[at(g, i, 1)]);
        x[at(g, i, ny + 1)] = sgn(b == B_NEGY, x[at(g, i, ny)]);
    }
    x[at(g, 0, 0)] = fmul(0.5f, fadd(x[at(g, 1, 0)], x[at(g, 0, 1)]));                            // lib.rs:40
    x[at(g, 0, ny + 1)] = fmul(0.5f, fadd(x[at(g, 1, ny + 1)], x[at(g, 0, ny)]));                 // lib.rs:41
    x[at(g, nx + 1, 0)] = fmul(0.5f, fadd(x[at(g, nx, 0)], x[at(g, nx + 1, 1)]));                 // lib.rs:42
    x[at(g, nx + 1, ny + 1)] = fmul(0.5f, fadd(x[at(g, nx, ny + 1)], x[at(g, nx + 1, ny)]));      // lib.rs:43
}

__global__ void k_serial_set_bnd(float* x, Geom g, int b) { serial_set_bnd(x, g, b); }

// lin_solve with an empty interior (lib.rs:52-70): each "sweep" is swap + set_bnd.
__global__ void k_serial_lin_solve(float* x, const float* x0, float* scratch, Geom g, float a, float c, int b,
                                   int iters) {
    (void)x0; (void)a; (void)c;
    for (int it = 0; it < iters; ++it) {
        float* t = scratch; scratch = x; x = t;
        serial_set_bnd(x, g, b);
    }
}

// project with an empty interior (lib.rs:146-169): only the six set_bnd calls act.
__global__ void k_serial_project(float* u, float* v, float* p, float* div, float* scratch, Geom g, int iters) {
    serial_set_bnd(div, g, B_POSITIVE);  // lib.rs:159
    serial_set_bnd(p, g, B_POSITIVE);    // lib.rs:160
    float* x = p;
    float* s = scratch;
    for (int it = 0; it < iters; ++it) {  // lib.rs:162
        float* t = s; s = x; x = t;
        serial_set_bnd(x, g, B_POSITIVE);
    }
    serial_set_bnd(u, g, B_NEGX);  // lib.rs:167
    serial_set_bnd(v, g, B_NEGY);  // lib.rs:168
}

}  // namespace fruid
