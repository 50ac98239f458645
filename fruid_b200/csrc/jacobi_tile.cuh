// jacobi_tile.cuh -- T fused Jacobi sweeps of lin_solve (src/lib.rs:46-71) per launch.
//
// Register-tile temporal blocking.  A CTA owns a tile of 128 columns x (R*NW) rows; warp
// `wi` owns R consecutive rows, lane `l` owns 4 consecutive columns (one float4 per row).
// The iterate x AND the right-hand side x0 live in registers for all T sweeps; x-neighbours
// travel by warp shuffle, the two rows a warp needs from its neighbours go through a
// double-buffered shared-memory exchange (one __syncthreads per sweep).  HBM traffic per
// launch is one read of x, one read of x0 and one write of the result (12 B/cell times the
// halo overhead) instead of 12 B/cell per sweep.
//
// Tiles overlap by 2H cells (H = T rounded up to even, so tile origins stay float4
// aligned): after s sweeps the outermost s rings of a tile are stale, and only the cells
// at least H away from every tile edge that is not a domain edge are stored.
//
// set_bnd (src/lib.rs:69, run after every sweep) is folded in by "mirror on read": border
// cells are never maintained during the sweeps; from sweep 2 on, a read of a border cell is
// replaced by sigma * (adjacent interior cell of the same iterate), which is exactly what
// set_bnd wrote there (lib.rs:31-37; negation is exact).  Sweep 1 reads the incoming border
// as loaded (lib.rs:56-59 read x before any set_bnd).  The 5-point stencil never reads a
// corner; border rows/columns and the four corners (lib.rs:40-43) of the final iterate are
// produced at store time by the thread that owns the adjacent interior cell.
//
// Arithmetic: sum = ((l + r) + u) + d; out = (x0 + a*sum) / c  (lib.rs:61-65), built from
// __fadd_rn/__fmul_rn/__fdiv_rn only.  MODE_POW2 replaces the division by a multiplication
// with 1/c when c is a power of two (bit-identical: same real value, same rounding);
// MODE_ONE drops it when c == 1.  a*sum is always computed (0*inf, 1*NaN etc. must match).
#pragma once
#include "common.cuh"

namespace fruid {

enum : int { MODE_GENERAL = 0, MODE_POW2 = 1, MODE_ONE = 2 };

struct TileArgs {
    float* out;
    const float* x;
    const float* x0;
    Geom g;
    float a, c;  // c holds 1/c in MODE_POW2
    int b;       // bounds kind
    int T;       // sweeps in this launch
    int H;       // halo = T rounded up to even
    int SX, SY;  // tile strides = extent - 2H
    int ntx, nty;
};

template <int MODE>
__device__ __forceinline__ float jac_upd(float z, float l, float r, float u, float d, float a, float c) {
    const float sum = fadd(fadd(fadd(l, r), u), d);
    const float v = fadd(z, fmul(a, sum));
    if (MODE == MODE_GENERAL) return fdiv(v, c);
    if (MODE == MODE_POW2) return fmul(v, c);
    return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 sgn4(bool neg, float4 v) {
    return neg ? make_float4(-v.x, -v.y, -v.z, -v.w) : v;
}
__device__ __forceinline__ float comp(const float4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

template <int R, int NW, int MODE, bool EDGE>
__device__ __forceinline__ void jacobi_tile_body(const TileArgs& A, float4 (*halo)[NW][2][32]) {
    constexpr int TH = R * NW;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const Geom& g = A.g;
    const int kx = blockIdx.x, ky = blockIdx.y;
    const int ox = kx * A.SX, oy = ky * A.SY;
    const int gx0 = ox + 4 * lane;  // global column of component .x
    const int gyw = oy + wi * R;    // global row of this warp's row 0
    const bool negx = (A.b == B_NEGX), negy = (A.b == B_NEGY);
    const float a = A.a, c = A.c;

    float4 xr[R], zr[R];
    const bool colok = gx0 < (int)g.pitch;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int gy = gyw + r;
        if (colok && gy < g.h) {
            const size_t k = (size_t)gy * g.pitch + gx0;
            xr[r] = ldg4(A.x + k);
            zr[r] = ldg4(A.x0 + k);
        } else {
            xr[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            zr[r] = xr[r];
        }
    }

    // EDGE tiles: which of my components sit next to a domain border column
    // (gx == 1: left neighbour is border column 0; gx == w-2: right neighbour is column w-1).
    bool l1 = false, r0 = false, r1 = false, r2 = false, r3 = false;
    if (EDGE) {
        l1 = (gx0 + 1 == 1);  // ox == 0 && lane == 0 (tile origins are multiples of 4)
        r0 = (gx0 + 0 == g.w - 2); r1 = (gx0 + 1 == g.w - 2); r2 = (gx0 + 2 == g.w - 2); r3 = (gx0 + 3 == g.w - 2);
    }

    for (int s = 1; s <= A.T; ++s) {
        const int p = s & 1;
        halo[p][wi][0][lane] = xr[0];
        halo[p][wi][1][lane] = xr[R - 1];
        __syncthreads();
        const float4 htop = (wi > 0) ? halo[p][wi - 1][1][lane] : xr[0];
        const float4 hbot = (wi < NW - 1) ? halo[p][wi + 1][0][lane] : xr[R - 1];
        const bool sub = EDGE && (s > 1);
        float4 prev = htop;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 cc = xr[r];
            float4 up = prev;
            float4 dn = (r == R - 1) ? hbot : xr[r + 1];
            float lft = __shfl_up_sync(FULL, cc.w, 1);
            float rgt = __shfl_down_sync(FULL, cc.x, 1);
            float ly = cc.x, rx = cc.y, ry = cc.z, rz = cc.w;
            if (EDGE) {
                const int gy = gyw + r;
                if (sub && gy == 1) up = sgn4(negy, cc);           // row 0 = sy * row 1       (lib.rs:36)
                if (sub && gy == g.h - 2) dn = sgn4(negy, cc);     // row ny+1 = sy * row ny   (lib.rs:37)
                if (sub && l1) ly = sgn(negx, cc.y);               // col 0 = sx * col 1       (lib.rs:31)
                if (sub && r0) rx = sgn(negx, cc.x);               // col nx+1 = sx * col nx   (lib.rs:32)
                if (sub && r1) ry = sgn(negx, cc.y);
                if (sub && r2) rz = sgn(negx, cc.z);
                if (sub && r3) rgt = sgn(negx, cc.w);
            }
            float4 n;
            n.x = jac_upd<MODE>(zr[r].x, lft, rx, up.x, dn.x, a, c);
            n.y = jac_upd<MODE>(zr[r].y, ly, ry, up.y, dn.y, a, c);
            n.z = jac_upd<MODE>(zr[r].z, cc.y, rz, up.z, dn.z, a, c);
            n.w = jac_upd<MODE>(zr[r].w, cc.z, rgt, up.w, dn.w, a, c);
            prev = cc;
            xr[r] = n;
        }
    }

    // ---- store the cells that are H away from every non-domain tile edge -----------------
    const bool lastx = (kx == A.ntx - 1), lasty = (ky == A.nty - 1);
    const int xlo = (kx == 0) ? 0 : ox + A.H, xhi = lastx ? g.w : ox + 128 - A.H;   // [xlo, xhi)
    const int ylo = (ky == 0) ? 1 : oy + A.H, yhi = lasty ? g.h - 1 : oy + TH - A.H;  // interior rows only
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int gy = gyw + r;
        if (gy < ylo || gy >= yhi) continue;  // warp-uniform
        float4 f = xr[r];
        float vleft = 0.f;  // value at gx0-1 (only needed when my .x is the right border column)
        if (EDGE) {
            vleft = __shfl_up_sync(FULL, f.w, 1);
            // column borders of this interior row (lib.rs:31-32)
            if (gx0 == 0) f.x = sgn(negx, f.y);
            if (gx0 + 0 == g.w - 1) f.x = sgn(negx, vleft);
            if (gx0 + 1 == g.w - 1) f.y = sgn(negx, f.x);
            if (gx0 + 2 == g.w - 1) f.z = sgn(negx, f.y);
            if (gx0 + 3 == g.w - 1) f.w = sgn(negx, f.z);
        }
        float* row = A.out + (size_t)gy * g.pitch;
        const bool v0 = gx0 + 0 >= xlo && gx0 + 0 < xhi, v1 = gx0 + 1 >= xlo && gx0 + 1 < xhi;
        const bool v2 = gx0 + 2 >= xlo && gx0 + 2 < xhi, v3 = gx0 + 3 >= xlo && gx0 + 3 < xhi;
        if (v0 && v1 && v2 && v3) {
            *reinterpret_cast<float4*>(row + gx0) = f;
        } else {
            if (v0) row[gx0 + 0] = f.x;
            if (v1) row[gx0 + 1] = f.y;
            if (v2) row[gx0 + 2] = f.z;
            if (v3) row[gx0 + 3] = f.w;
        }
        if (EDGE && (gy == 1 || gy == g.h - 2)) {
            // border row(s) mirrored from this row (lib.rs:36-37) and their corners (lib.rs:40-43):
            // corner = 0.5*(row-edge + col-edge) = 0.5*(sy*v + sx*v), v = diagonal interior cell;
            // f already holds sx*v in the border columns.
            float4 e = sgn4(negy, f);
            if (gx0 == 0) e.x = fmul(0.5f, fadd(sgn(negy, f.y), f.x));
            if (gx0 + 0 == g.w - 1) e.x = fmul(0.5f, fadd(sgn(negy, vleft), f.x));
            if (gx0 + 1 == g.w - 1) e.y = fmul(0.5f, fadd(sgn(negy, f.x), f.y));
            if (gx0 + 2 == g.w - 1) e.z = fmul(0.5f, fadd(sgn(negy, f.y), f.z));
            if (gx0 + 3 == g.w - 1) e.w = fmul(0.5f, fadd(sgn(negy, f.z), f.w));
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (side == 0 && gy != 1) continue;
                if (side == 1 && gy != g.h - 2) continue;
                float* brow = A.out + (size_t)(side == 0 ? 0 : g.h - 1) * g.pitch;
                if (v0) brow[gx0 + 0] = e.x;
                if (v1) brow[gx0 + 1] = e.y;
                if (v2) brow[gx0 + 2] = e.z;
                if (v3) brow[gx0 + 3] = e.w;
            }
        }
    }
}

template <int R, int NW, int MODE>
__global__ void __launch_bounds__(NW * 32) k_jacobi_tile(TileArgs A) {
    __shared__ float4 halo[2][NW][2][32];
    const int ox = blockIdx.x * A.SX, oy = blockIdx.y * A.SY;
    const bool edge = (blockIdx.x == 0) || (blockIdx.y == 0) || (ox + 128 >= A.g.w - 1) || (oy + R * NW >= A.g.h - 1);
    if (edge)
        jacobi_tile_body<R, NW, MODE, true>(A, halo);
    else
        jacobi_tile_body<R, NW, MODE, false>(A, halo);
}

// ---- host side -----------------------------------------------------------------------------
struct TileCfg { int R, NW; };
static TileCfg g_tile_cfg = {8, 8};

static inline int jacobi_tile_max_T(int R, int NW) {
    const int th = R * NW;
    int H = (th < 128 ? th : 128) / 2 - 2;  // keep at least 4 output rows/cols per tile
    return H & ~1;
}

static inline int jacobi_tile_default_fused(const Geom& g, int iters) {
    (void)g;
    int T = 10;
    const int mx = jacobi_tile_max_T(g_tile_cfg.R, g_tile_cfg.NW);
    if (T > mx) T = mx;
    if (T > iters) T = iters;
    return T < 1 ? 1 : T;
}

template <int R, int NW>
static int jacobi_tile_launch_cfg(const TileArgs& A, int mode, cudaStream_t st) {
    dim3 grid(A.ntx, A.nty), block(NW * 32);
    if (mode == MODE_GENERAL) k_jacobi_tile<R, NW, MODE_GENERAL><<<grid, block, 0, st>>>(A);
    else if (mode == MODE_POW2) k_jacobi_tile<R, NW, MODE_POW2><<<grid, block, 0, st>>>(A);
    else k_jacobi_tile<R, NW, MODE_ONE><<<grid, block, 0, st>>>(A);
    return 0;
}

static inline int tiles_needed(int extent, int tile, int stride) {
    if (extent <= tile) return 1;
    return (extent - tile + stride - 1) / stride + 1;
}

// Returns 0 or a negative fruid_status (-7 = bad argument).
static int jacobi_tile_launch(float* out, const float* x, const float* x0, const Geom& g, float a, float c, int b, int T,
                              cudaStream_t st) {
    const int R = g_tile_cfg.R, NW = g_tile_cfg.NW, TH = R * NW;
    if (T < 1 || T > jacobi_tile_max_T(R, NW)) return -7;
    TileArgs A;
    A.out = out; A.x = x; A.x0 = x0; A.g = g; A.a = a; A.b = b; A.T = T;
    A.H = T + (T & 1);
    A.SX = 128 - 2 * A.H;
    A.SY = TH - 2 * A.H;
    A.ntx = tiles_needed(g.w, 128, A.SX);
    A.nty = tiles_needed(g.h, TH, A.SY);
    int mode = MODE_GENERAL;
    A.c = c;
    if (c == 1.0f) {
        mode = MODE_ONE;
    } else {
        int e = 0;
        const float m = frexpf(c, &e);
        if (m == 0.5f && e > -100 && e < 100) { mode = MODE_POW2; A.c = 1.0f / c; }  // exact reciprocal
    }
#define FRUID_TILE_CASE(r, nw) if (R == r && NW == nw) return jacobi_tile_launch_cfg<r, nw>(A, mode, st)
    FRUID_TILE_CASE(8, 8);
    FRUID_TILE_CASE(8, 16);
    FRUID_TILE_CASE(12, 8);
    FRUID_TILE_CASE(12, 16);
    FRUID_TILE_CASE(16, 8);
    FRUID_TILE_CASE(16, 12);
    FRUID_TILE_CASE(6, 16);
    FRUID_TILE_CASE(4, 16);
#undef FRUID_TILE_CASE
    return -7;
}

static inline int jacobi_tile_set_cfg(int R, int NW) {
    const int ok[][2] = {{8, 8}, {8, 16}, {12, 8}, {12, 16}, {16, 8}, {16, 12}, {6, 16}, {4, 16}};
    for (auto& p : ok)
        if (p[0] == R && p[1] == NW) { g_tile_cfg = {R, NW}; return 0; }
    return -7;
}

}  // namespace fruid
