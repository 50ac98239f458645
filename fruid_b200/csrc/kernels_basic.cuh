// kernels_basic.cuh -- one-pass kernels: add_source, single-sweep Jacobi, divergence,
// gradient-subtract, advection, set_bnd.  One thread per interior cell, x fastest so a
// warp reads/writes one contiguous 128-byte row segment.  Each kernel cites the
// reference loop nest it replaces (src/lib.rs).
#pragma once
#include "common.cuh"

namespace fruid {

// add_source, src/lib.rs:5-10: x[k] += s[k]*dt over ALL w*h cells.  Two fields per
// launch (the u and v calls at lib.rs:190-191); pass x2 == nullptr for one.
__global__ void k_add_source(float* __restrict__ x1, const float* __restrict__ s1, float* __restrict__ x2,
                             const float* __restrict__ s2, Geom g, float dt) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.w || y >= g.h) return;
    const size_t k = at(g, x, y);
    x1[k] = fadd(x1[k], fmul(s1[k], dt));
    if (x2) x2[k] = fadd(x2[k], fmul(s2[k], dt));
}

// set_bnd as a stand-alone operator, src/lib.rs:27-44 (only the per-operator ABI uses
// it; inside the step it is always folded into the producing kernel).
__global__ void k_set_bnd(float* x, Geom g, int b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > g.nx || j > g.ny) return;
    if (i != 1 && i != g.nx && j != 1 && j != g.ny) return;
    store_with_bnd(x, g, b, i, j, x[at(g, i, j)]);
}

// One Jacobi sweep of lin_solve (src/lib.rs:53-67) + the set_bnd of lib.rs:69.
//   out[i,j] = (x0[i,j] + a*(((x[i-1,j] + x[i+1,j]) + x[i,j-1]) + x[i,j+1])) / c
__global__ void k_jacobi1(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ x0, Geom g,
                          float a, float c, int b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > g.nx || j > g.ny) return;
    const float left = x[at(g, i - 1, j)], right = x[at(g, i + 1, j)];
    const float up = x[at(g, i, j - 1)], down = x[at(g, i, j + 1)];
    const float sum = fadd(fadd(fadd(left, right), up), down);
    const float v = fdiv(fadd(x0[at(g, i, j)], fmul(a, sum)), c);
    store_with_bnd(out, g, b, i, j, v);
}

// First half of project (src/lib.rs:149-160): zero div and p, two diff_acc passes into
// div, set_bnd(Positive) on both.  p ends up +0.0 everywhere (border and corners too).
//   div = (0 + sx*(u[i+1,j]-u[i-1,j])) + sy*(v[i,j+1]-v[i,j-1])
__global__ void k_divergence(float* __restrict__ div, float* __restrict__ p, const float* __restrict__ u,
                             const float* __restrict__ v, Geom g, float sx, float sy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > g.nx || j > g.ny) return;
    const float dx = fsub(u[at(g, i + 1, j)], u[at(g, i - 1, j)]);
    const float dy = fsub(v[at(g, i, j + 1)], v[at(g, i, j - 1)]);
    const float d = fadd(fadd(0.0f, fmul(sx, dx)), fmul(sy, dy));
    store_with_bnd(div, g, B_POSITIVE, i, j, d);
    if (p) store_with_bnd(p, g, B_POSITIVE, i, j, 0.0f);
}

// Second half of project (src/lib.rs:164-168): u += gx*(p[i+1,j]-p[i-1,j]),
// v += gy*(p[i,j+1]-p[i,j-1]), then set_bnd(NegX,u), set_bnd(NegY,v).  In place: each
// thread touches only its own cell of u,v (+ the border cells it owns).
__global__ void k_gradient(float* __restrict__ u, float* __restrict__ v, const float* __restrict__ p, Geom g,
                           float gx, float gy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > g.nx || j > g.ny) return;
    const float dx = fsub(p[at(g, i + 1, j)], p[at(g, i - 1, j)]);
    const float dy = fsub(p[at(g, i, j + 1)], p[at(g, i, j - 1)]);
    const size_t k = at(g, i, j);
    store_with_bnd(u, g, B_NEGX, i, j, fadd(u[k], fmul(gx, dx)));
    store_with_bnd(v, g, B_NEGY, i, j, fadd(v[k], fmul(gy, dy)));
}

// Back-trace of one cell, src/lib.rs:92-116.  Returns the four gather offsets and the
// two bilinear weights.  `x as usize` on the clamped value = cvt.rzi (NaN -> 0).
struct Trace {
    int i0, j0;
    float s1, t1;
};
__device__ __forceinline__ Trace backtrace(int i, int j, float uc, float vc, float dt0, float dt1, float xmax,
                                           float ymax) {
    float x = fsub((float)i, fmul(dt0, uc));
    float y = fsub((float)j, fmul(dt1, vc));
    if (x < 0.5f) x = 0.5f;
    if (x > xmax) x = xmax;
    if (y < 0.5f) y = 0.5f;
    if (y > ymax) y = ymax;
    Trace t;
    t.i0 = (int)__float2uint_rz(x);
    t.j0 = (int)__float2uint_rz(y);
    t.s1 = fsub(x, (float)t.i0);
    t.t1 = fsub(y, (float)t.j0);
    return t;
}
__device__ __forceinline__ float gather(const float* __restrict__ d0, const Geom& g, const Trace& t) {
    const size_t k00 = at(g, t.i0, t.j0);
    const float a00 = __ldg(d0 + k00), a01 = __ldg(d0 + k00 + g.pitch);
    const float a10 = __ldg(d0 + k00 + 1), a11 = __ldg(d0 + k00 + 1 + g.pitch);
    return mixf(mixf(a00, a01, t.t1), mixf(a10, a11, t.t1), t.s1);
}

// advect, src/lib.rs:84-126 (+ set_bnd at :125), for NF fields sharing one back-trace:
// NF = 2 is the velocity self-advection of lib.rs:204-205 (d0 aliases u / v), NF = 1 the
// density advection of lib.rs:178.
// NF = 3, 4: several dye fields advected by one velocity (the four DensitySims of main.rs:114-118).
constexpr int ADVECT_MAX_FIELDS = 4;
struct AdvectArgs {
    float* d[ADVECT_MAX_FIELDS];
    const float* d0[ADVECT_MAX_FIELDS];
    int b[ADVECT_MAX_FIELDS];
};
template <int NF>
__global__ void k_advect(AdvectArgs a, const float* __restrict__ u, const float* __restrict__ v, Geom g, float dt0,
                         float dt1, float xmax, float ymax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > g.nx || j > g.ny) return;
    const size_t k = at(g, i, j);
    const Trace t = backtrace(i, j, __ldg(u + k), __ldg(v + k), dt0, dt1, xmax, ymax);
#pragma unroll
    for (int f = 0; f < NF; ++f) store_with_bnd(a.d[f], g, a.b[f], i, j, gather(a.d0[f], g, t));
}

}  // namespace fruid

namespace fruid {
// `data_mut().fill(v)`, src/main.rs:105-108.
__global__ void k_fill(float* __restrict__ x, Geom g, float val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < g.w && j < g.h) x[at(g, i, j)] = val;
}
}  // namespace fruid
