// kernels_vec.cuh -- float4-per-thread versions of the single-pass operators (add_source,
// divergence, gradient-subtract, advection).  One thread owns four consecutive cells of one
// row, so every global access is a 16-byte vector and a warp covers one 512-byte row
// segment.  Boundary cells are written by the thread that owns the adjacent interior cell
// (set_bnd folded into the producer, src/lib.rs:27-44).  Requires nx >= 1, ny >= 1,
// pitch % 4 == 0 and 16-byte aligned bases (true for every library-owned field).
#pragma once
#include "common.cuh"
#include "kernels_basic.cuh"

namespace fruid {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 neg4(bool neg, float4 v) { return neg ? make_float4(-v.x, -v.y, -v.z, -v.w) : v; }

// Stores the four new values `val` of interior row gy (components whose column is a border or
// lies outside the field are ignored) together with every border cell they determine:
//   x[0,j] = sx*x[1,j], x[nx+1,j] = sx*x[nx,j]            lib.rs:31-32
//   x[i,0] = sy*x[i,1], x[i,ny+1] = sy*x[i,ny]            lib.rs:36-37
//   corner = 0.5*(row-edge + col-edge)                    lib.rs:40-43
// `ly` is the LOCAL row (its global row is ly + g.gy0); rows ly-1 / ly+1 must exist locally.
__device__ __forceinline__ void store_row4_bnd(float* __restrict__ out, const Geom& g, int b, int gx0, int ly, float4 val) {
    const int gy = ly + g.gy0;
    const bool negx = (b == B_NEGX), negy = (b == B_NEGY);
    const int kr = g.w - 1 - gx0;  // component index of the right border column (0..3 if inside this float4)
    if (gx0 == 0) val.x = sgn(negx, val.y);
    if (kr == 1) val.y = sgn(negx, val.x);
    if (kr == 2) val.z = sgn(negx, val.y);
    if (kr == 3) val.w = sgn(negx, val.z);
    const float extra = sgn(negx, val.w);  // border value at gx0+4 when kr == 4 (I own column w-2 in .w)
    // kr == 0: my .x is the right border but its source is the previous thread's .w -> it writes it.
    const bool w0 = (kr != 0) && (gx0 + 0 < g.w), w1 = gx0 + 1 < g.w, w2 = gx0 + 2 < g.w, w3 = gx0 + 3 < g.w;
    float* row = out + (size_t)ly * g.pitch + gx0;
    if (w0 && w3) {
        *reinterpret_cast<float4*>(row) = val;
    } else {
        if (w0) row[0] = val.x;
        if (w1) row[1] = val.y;
        if (w2) row[2] = val.z;
        if (w3) row[3] = val.w;
    }
    if (kr == 4) row[4] = extra;
    const bool top = (gy == 1), bot = (gy == g.ny);
    if (top | bot) {
        float4 e = neg4(negy, val);
        if (gx0 == 0) e.x = fmul(0.5f, fadd(sgn(negy, val.y), val.x));
        if (kr == 1) e.y = fmul(0.5f, fadd(sgn(negy, val.x), val.y));
        if (kr == 2) e.z = fmul(0.5f, fadd(sgn(negy, val.y), val.z));
        if (kr == 3) e.w = fmul(0.5f, fadd(sgn(negy, val.z), val.w));
        const float ecorner = fmul(0.5f, fadd(sgn(negy, val.w), extra));
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            if (side == 0 ? !top : !bot) continue;
            float* brow = out + (size_t)(side == 0 ? ly - 1 : ly + 1) * g.pitch + gx0;
            if (w0 && w3) {
                *reinterpret_cast<float4*>(brow) = e;
            } else {
                if (w0) brow[0] = e.x;
                if (w1) brow[1] = e.y;
                if (w2) brow[2] = e.z;
                if (w3) brow[3] = e.w;
            }
            if (kr == 4) brow[4] = ecorner;
        }
    }
}

// add_source over all cells (src/lib.rs:5-10), two fields per launch; rows as float4.
__global__ void k_add_source4(float* __restrict__ x1, const float* __restrict__ s1, float* __restrict__ x2,
                              const float* __restrict__ s2, size_t n4, float dt) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4* X1 = reinterpret_cast<float4*>(x1);
    const float4 a = X1[i], s = ld4(s1 + 4 * i);
    X1[i] = make_float4(fadd(a.x, fmul(s.x, dt)), fadd(a.y, fmul(s.y, dt)), fadd(a.z, fmul(s.z, dt)), fadd(a.w, fmul(s.w, dt)));
    if (x2) {
        float4* X2 = reinterpret_cast<float4*>(x2);
        const float4 c = X2[i], t = ld4(s2 + 4 * i);
        X2[i] = make_float4(fadd(c.x, fmul(t.x, dt)), fadd(c.y, fmul(t.y, dt)), fadd(c.z, fmul(t.z, dt)), fadd(c.w, fmul(t.w, dt)));
    }
}

// Left/right neighbours of a float4 at (gx0, gy): values at gx0-1 and gx0+4 (0 outside the row).
__device__ __forceinline__ void edge_vals(const float* __restrict__ f, const Geom& g, int gx0, size_t k, float& lm, float& rp) {
    lm = (gx0 > 0) ? __ldg(f + k - 1) : 0.0f;
    rp = (gx0 + 4 < (int)g.pitch) ? __ldg(f + k + 4) : 0.0f;
}

// First half of project (src/lib.rs:149-160): div = (0 + sx*(u[i+1,j]-u[i-1,j])) + sy*(v[i,j+1]-v[i,j-1]),
// set_bnd(Positive, div); p := +0 everywhere (skipped when p == nullptr: the pressure solve
// is then told that its initial iterate is identically zero).
__global__ void k_divergence4(float* __restrict__ div, float* __restrict__ p, const float* __restrict__ u,
                              const float* __restrict__ v, Geom g, float sx, float sy) {
    const int gx0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int gy = blockIdx.y * blockDim.y + threadIdx.y + 1;  // LOCAL row; needs rows gy-1, gy+1
    if (gx0 >= g.w || gy > g.h - 2) return;
    const size_t k = (size_t)gy * g.pitch + gx0;
    const float4 uc = ld4(u + k), vu = ld4(v + k - g.pitch), vd = ld4(v + k + g.pitch);
    float ul, ur;
    edge_vals(u, g, gx0, k, ul, ur);
    float4 d;
    d.x = fadd(fadd(0.0f, fmul(sx, fsub(uc.y, ul))), fmul(sy, fsub(vd.x, vu.x)));
    d.y = fadd(fadd(0.0f, fmul(sx, fsub(uc.z, uc.x))), fmul(sy, fsub(vd.y, vu.y)));
    d.z = fadd(fadd(0.0f, fmul(sx, fsub(uc.w, uc.y))), fmul(sy, fsub(vd.z, vu.z)));
    d.w = fadd(fadd(0.0f, fmul(sx, fsub(ur, uc.z))), fmul(sy, fsub(vd.w, vu.w)));
    store_row4_bnd(div, g, B_POSITIVE, gx0, gy, d);
    if (p) store_row4_bnd(p, g, B_POSITIVE, gx0, gy, make_float4(0.f, 0.f, 0.f, 0.f));
}

// Second half of project (src/lib.rs:164-168): u += gx*(p[i+1,j]-p[i-1,j]); v += gy*(p[i,j+1]-p[i,j-1]);
// set_bnd(NegX,u); set_bnd(NegY,v).  In place; each thread touches only its own cells.
__global__ void k_gradient4(float* __restrict__ u, float* __restrict__ v, const float* __restrict__ p, Geom g, float gxs,
                            float gys) {
    const int gx0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int gy = blockIdx.y * blockDim.y + threadIdx.y + 1;  // LOCAL row; needs rows gy-1, gy+1
    if (gx0 >= g.w || gy > g.h - 2) return;
    const size_t k = (size_t)gy * g.pitch + gx0;
    const float4 pc = ld4(p + k), pu = ld4(p + k - g.pitch), pd = ld4(p + k + g.pitch);
    float pl, pr;
    edge_vals(p, g, gx0, k, pl, pr);
    const float4 uo = *reinterpret_cast<const float4*>(u + k), vo = *reinterpret_cast<const float4*>(v + k);
    float4 un, vn;
    un.x = fadd(uo.x, fmul(gxs, fsub(pc.y, pl)));
    un.y = fadd(uo.y, fmul(gxs, fsub(pc.z, pc.x)));
    un.z = fadd(uo.z, fmul(gxs, fsub(pc.w, pc.y)));
    un.w = fadd(uo.w, fmul(gxs, fsub(pr, pc.z)));
    vn.x = fadd(vo.x, fmul(gys, fsub(pd.x, pu.x)));
    vn.y = fadd(vo.y, fmul(gys, fsub(pd.y, pu.y)));
    vn.z = fadd(vo.z, fmul(gys, fsub(pd.z, pu.z)));
    vn.w = fadd(vo.w, fmul(gys, fsub(pd.w, pu.w)));
    store_row4_bnd(u, g, B_NEGX, gx0, gy, un);
    store_row4_bnd(v, g, B_NEGY, gx0, gy, vn);
}

// Where the rows of a field live when it is y-slab decomposed over several GPUs: rank r holds
// global rows [row0[r], row0[r] + rows[r]) ... owned rows [own0[r], own0[r+1]) at base[f][r]
// (a peer-mapped pointer, NVLink).  A gather tap at global row gj reads the copy of its OWNER.
struct SlabTable {
    int world;                 // 1 = single GPU (base[.][0] = local field, row0[0] = 0)
    int me;                    // my rank
    int my_lo, my_hi, my_row0; // my owned global rows [my_lo, my_hi) and the global row of my local row 0:
                               // taps inside them use the local field pointer (no table lookup)
    int own0[9];               // owned global rows of rank r: [own0[r], own0[r+1])
    int row0[8];               // global row of local row 0 of rank r
    const float* base[2][8];   // [field][rank]
};
__device__ __noinline__ const float* slab_row_remote(const SlabTable& t, int f, int gj, size_t pitch) {
    int r = 0;
    while (r + 1 < t.world && gj >= t.own0[r + 1]) ++r;
    return t.base[f][r] + (size_t)(gj - t.row0[r]) * pitch;
}
__device__ __forceinline__ const float* slab_row(const SlabTable& t, int f, const float* local, int gj, size_t pitch) {
    if (gj >= t.my_lo && gj < t.my_hi) return local + (size_t)(gj - t.my_row0) * pitch;
    return slab_row_remote(t, f, gj, pitch);  // rare: the tap belongs to another rank (peer memory)
}
template <bool MULTI>
__device__ __forceinline__ float gather_slab(const SlabTable& tab, int f, const float* __restrict__ d0, const Geom& g,
                                             const Trace& t) {
    float a00, a01, a10, a11;
    if (MULTI) {
        const float* r0 = slab_row(tab, f, d0, t.j0, g.pitch) + t.i0;
        const float* r1 = slab_row(tab, f, d0, t.j0 + 1, g.pitch) + t.i0;
        a00 = __ldg(r0); a10 = __ldg(r0 + 1); a01 = __ldg(r1); a11 = __ldg(r1 + 1);
    } else {
        const float* r0 = d0 + (size_t)t.j0 * g.pitch + t.i0;
        a00 = __ldg(r0); a10 = __ldg(r0 + 1); a01 = __ldg(r0 + g.pitch); a11 = __ldg(r0 + g.pitch + 1);
    }
    return mixf(mixf(a00, a01, t.t1), mixf(a10, a11, t.t1), t.s1);
}

// advect (src/lib.rs:84-126) for NF fields sharing one back-trace, four cells per thread.
// Output rows: local rows [ly0, ly1) (all interior rows on one GPU; the owned rows of a slab).
template <int NF, bool MULTI>
__global__ void k_advect4(AdvectArgs a, const float* __restrict__ u, const float* __restrict__ v, Geom g, float dt0,
                          float dt1, float xmax, float ymax, int ly0, int ly1, const __grid_constant__ SlabTable tab) {
    const int gx0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int ly = blockIdx.y * blockDim.y + threadIdx.y + ly0;
    if (gx0 >= g.w || ly >= ly1) return;
    const int gy = ly + g.gy0;  // global row: the back-trace and the gather use global coordinates
    const size_t k = (size_t)ly * g.pitch + gx0;
    const float4 uc = ld4(u + k), vc = ld4(v + k);
    Trace t[4];
    t[0] = backtrace(gx0 + 0, gy, uc.x, vc.x, dt0, dt1, xmax, ymax);
    t[1] = backtrace(gx0 + 1, gy, uc.y, vc.y, dt0, dt1, xmax, ymax);
    t[2] = backtrace(gx0 + 2, gy, uc.z, vc.z, dt0, dt1, xmax, ymax);
    t[3] = backtrace(gx0 + 3, gy, uc.w, vc.w, dt0, dt1, xmax, ymax);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        float4 r;
        r.x = gather_slab<MULTI>(tab, f, a.d0[f], g, t[0]);
        r.y = gather_slab<MULTI>(tab, f, a.d0[f], g, t[1]);
        r.z = gather_slab<MULTI>(tab, f, a.d0[f], g, t[2]);
        r.w = gather_slab<MULTI>(tab, f, a.d0[f], g, t[3]);
        store_row4_bnd(a.d[f], g, a.b[f], gx0, ly, r);
    }
}

}  // namespace fruid
