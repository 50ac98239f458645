// jacobi_tile.cuh -- T fused Jacobi sweeps of lin_solve (src/lib.rs:46-71) per launch.
//
// Register-tile temporal blocking for sm_100a.  A CTA owns a tile of 128 columns x (R*NW)
// rows; warp `wi` owns R consecutive rows, lane `l` owns 4 consecutive columns (one float4
// per row).  The kernel is PERSISTENT (one CTA per SM looping over tiles): while tile i is
// being swept, one elected thread has already issued the TMA loads (cp.async.bulk.tensor.2d,
// mbarrier completion, zero fill outside the field) of tile i+1's x and x0 into shared
// memory, so HBM latency and bandwidth are hidden behind the sweeps.  The iterate x lives in
// REGISTERS for all T sweeps and so does the right-hand side x0 (both copied once per tile
// from their TMA staging buffers), which keeps the shared-memory/shuffle (MIO) path free for
// what cannot live in registers.  x-neighbours
// travel by warp shuffle; the two rows a warp needs from its neighbour warps go through a
// double-buffered shared-memory exchange (one __syncthreads per sweep).  HBM traffic per
// launch is one read of x, one read of x0 and one write of the result (12 B/cell times the
// halo overhead) instead of 12 B/cell per sweep; the kernel is issue-bound, not HBM-bound.
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
// MODE_ONE drops it when c == 1; MODE_POISSON additionally drops the multiplication by
// a == 1 (1*sum == sum bit-for-bit).  Otherwise a*sum is always computed (0*inf etc.).
// MODE_RECIP divides by a general c without the ~25-instruction IEEE division sequence: with the
// reciprocal as a double-word (zh, zl), q = fma(v, zh, v*zl) is the correctly rounded v/c (division
// by a constant with an FMA, Brisebarre-Muller-Raina 2004) -- for "almost all" c; the claim is not
// taken on trust: before a value of c is used, a kernel compares q with __fdiv_rn(v, c) for ALL 2^23
// significands of v (results scale exactly with the exponent while nothing under/overflows), and
// tries MODE_RECIP3 if a single one differs (it does for about one divisor in a few hundred -- among them
// c = 1 + 4*(0.01*1e-6*4094*4094), the viscous 4096^2 bench case): q0 = v*zh, r = fma(-q0, c, v) (the exact
// residual), q = fma(r, zh, q0), Markstein's correction step, one instruction more; verified the same way,
// and only if that fails too MODE_GENERAL is used.  Inputs outside [2^-60, 2^60) (zeros, denormals, huge
// values, inf, NaN) take __fdiv_rn per element.
#pragma once
#include <cuda.h>

#include <cstring>
#include <iterator>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace fruid {

enum : int { MODE_GENERAL = 0, MODE_POW2 = 1, MODE_ONE = 2, MODE_POISSON = 3, MODE_RECIP = 4, MODE_RECIP3 = 5 };

struct TileArgs {
    CUtensorMap tm_x0;  // 2-D map over x0: dims (pitch, h) f32, box (128, R*NW), zero OOB fill
    CUtensorMap tm_x;   // same over the incoming iterate x
    float* out;
    const float* x;
    Geom g;
    float a, c;  // c holds 1/c in MODE_POW2 / MODE_POISSON
    float zh, zl;  // MODE_RECIP: 1/c as a double-word (zh = RN(1/c), zl = RN(1/c - zh))
    int b;       // bounds kind
    int T;       // sweeps in this launch
    int H;       // halo = T rounded up to even
    int SX, SY;  // tile strides = extent - 2H
    int ntx, nty;
    int x_is_zero;  // the incoming iterate is identically +0 (pressure solve, lib.rs:152,160): skip its load
    int* sched;     // [0] = next work item, [1] = CTAs finished (both self-reset by the last CTA)
    // add_source (lib.rs:5-10) folded into the first block of a diffuse: x0 += dt * x_initial
    // (the initial iterate of diffuse IS the source field, lib.rs:172-175,190-197).  The updated
    // right-hand side goes to z_out (a different buffer: other CTAs still read x0 as halo).
    int fuse_src;
    float src_dt;
    float* z_out;
};

// Work item -> tile.  Tiles touching the domain boundary execute more instructions, so they are
// handed out first (longest-first keeps the tail of the persistent grid short); interior tiles
// follow in row-major order so that concurrently running CTAs share halo rows/columns in L2.
__device__ __forceinline__ void tile_of_item(int item, int ntx, int nty, int& kx, int& ky) {
    if (ntx < 3 || nty < 3) { kx = item % ntx; ky = item / ntx; return; }
    if (item < ntx) { kx = item; ky = 0; return; }
    item -= ntx;
    if (item < ntx) { kx = item; ky = nty - 1; return; }
    item -= ntx;
    if (item < nty - 2) { kx = 0; ky = 1 + item; return; }
    item -= nty - 2;
    if (item < nty - 2) { kx = ntx - 1; ky = 1 + item; return; }
    item -= nty - 2;
    kx = 1 + item % (ntx - 2);
    ky = 1 + item / (ntx - 2);
}

// One float4 row of the update.  t = l + r per component (scalar adds: the operands are
// not register-pair aligned); the remaining adds and multiplies use the sm_100 packed
// f32x2 instructions (FADD2 / FMUL2: two IEEE round-to-nearest operations per issue slot,
// bit-identical to two scalar ones), which halves the issue pressure of this issue-bound
// kernel.  Order of operations is exactly lib.rs:61-65: ((l + r) + u) + d, then
// (x0 + a*sum) / c.
__device__ __forceinline__ bool recip_safe(float v) {  // 2^-60 <= |v| < 2^60
    return ((__float_as_uint(v) & 0x7f800000u) - (67u << 23)) < (120u << 23);
}
template <int MODE>
__device__ __forceinline__ float4 jac_row(const float4 z, float t0, float t1, float t2, float t3, const float4 up,
                                          const float4 dn, float a, float c, float zh = 0.f, float zl = 0.f) {
    float2 s01 = __fadd2_rn(make_float2(t0, t1), make_float2(up.x, up.y));
    float2 s23 = __fadd2_rn(make_float2(t2, t3), make_float2(up.z, up.w));
    s01 = __fadd2_rn(s01, make_float2(dn.x, dn.y));
    s23 = __fadd2_rn(s23, make_float2(dn.z, dn.w));
    if (MODE != MODE_POISSON) {
        // a*sum (for a == 1 the product is the sum itself, bit for bit).  SCALAR multiplies on
        // purpose: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under
        // -fmad=false, which would change the rounding (tests/test_abi_load.py greps for it).
        s01 = make_float2(fmul(a, s01.x), fmul(a, s01.y));
        s23 = make_float2(fmul(a, s23.x), fmul(a, s23.y));
    }
    s01 = __fadd2_rn(make_float2(z.x, z.y), s01);
    s23 = __fadd2_rn(make_float2(z.z, z.w), s23);
    if (MODE == MODE_GENERAL) return make_float4(fdiv(s01.x, c), fdiv(s01.y, c), fdiv(s23.x, c), fdiv(s23.y, c));
    if (MODE == MODE_RECIP || MODE == MODE_RECIP3) {
        // FMAs used as exact-arithmetic building blocks of the division, not contractions
        float2 q01, q23;
        if (MODE == MODE_RECIP) {
            q01 = __ffma2_rn(s01, make_float2(zh, zh), __fmul2_rn(s01, make_float2(zl, zl)));
            q23 = __ffma2_rn(s23, make_float2(zh, zh), __fmul2_rn(s23, make_float2(zl, zl)));
        } else {  // q0 = v*zh;  r = v - q0*c (exact);  q = q0 + r*zh
            const float2 p01 = __fmul2_rn(s01, make_float2(zh, zh)), p23 = __fmul2_rn(s23, make_float2(zh, zh));
            q01 = __ffma2_rn(__ffma2_rn(p01, make_float2(-c, -c), s01), make_float2(zh, zh), p01);
            q23 = __ffma2_rn(__ffma2_rn(p23, make_float2(-c, -c), s23), make_float2(zh, zh), p23);
        }
        // Range guard for the whole float4: min and max of the magnitudes (8 instructions instead of 12 integer ones;
        // fminf / fmaxf skip a NaN, whose quotient is a NaN on either path).
        const float mx = fmaxf(fmaxf(fabsf(s01.x), fabsf(s01.y)), fmaxf(fabsf(s23.x), fabsf(s23.y)));
        const float mn = fminf(fminf(fabsf(s01.x), fabsf(s01.y)), fminf(fabsf(s23.x), fabsf(s23.y)));
        if (mn >= 0x1p-60f && mx < 0x1p60f) return make_float4(q01.x, q01.y, q23.x, q23.y);
        const bool k0 = recip_safe(s01.x), k1 = recip_safe(s01.y), k2 = recip_safe(s23.x), k3 = recip_safe(s23.y);
        return make_float4(k0 ? q01.x : fdiv(s01.x, c), k1 ? q01.y : fdiv(s01.y, c), k2 ? q23.x : fdiv(s23.x, c),
                           k3 ? q23.y : fdiv(s23.y, c));
    }
    if (MODE == MODE_POW2 || MODE == MODE_POISSON) {
        s01 = __fmul2_rn(s01, make_float2(c, c));
        s23 = __fmul2_rn(s23, make_float2(c, c));
    }
    return make_float4(s01.x, s01.y, s23.x, s23.y);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 sgn4(bool neg, float4 v) {
    return neg ? make_float4(-v.x, -v.y, -v.z, -v.w) : v;
}

// ---- mbarrier / TMA primitives (PTX; SASS: SYNCS.*, UTMALDG) ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// EDGE tiles: flags saying which of my components sit next to a domain border column
// (gx == 1: left neighbour is border column 0; gx == w-2: right neighbour is column w-1).
struct EdgeFlags { bool l1, r0, r1, r2, r3, negx, negy; int h; float zh, zl; };

// New value of one float4 row from its old value cc, the rows above/below and x0 (z).
template <int MODE, bool EDGE>
__device__ __forceinline__ float4 jac_row_full(float4 cc, float4 up, float4 dn, const float4 z, float a, float c,
                                               const EdgeFlags& E, bool sub, int gy) {
    const unsigned FULL = 0xffffffffu;
    float lft = __shfl_up_sync(FULL, cc.w, 1);
    float rgt = __shfl_down_sync(FULL, cc.x, 1);
    float ly = cc.x, rx = cc.y, ry = cc.z, rz = cc.w;
    if (EDGE) {
        if (sub && gy == 1) up = sgn4(E.negy, cc);           // row 0 = sy * row 1       (lib.rs:36)
        if (sub && gy == E.h - 2) dn = sgn4(E.negy, cc);     // row ny+1 = sy * row ny   (lib.rs:37)
        if (sub && E.l1) ly = sgn(E.negx, cc.y);             // col 0 = sx * col 1       (lib.rs:31)
        if (sub && E.r0) rx = sgn(E.negx, cc.x);             // col nx+1 = sx * col nx   (lib.rs:32)
        if (sub && E.r1) ry = sgn(E.negx, cc.y);
        if (sub && E.r2) rz = sgn(E.negx, cc.z);
        if (sub && E.r3) rgt = sgn(E.negx, cc.w);
    }
    const float t0 = fadd(lft, rx), t1 = fadd(ly, ry), t2 = fadd(cc.y, rz), t3 = fadd(cc.z, rgt);  // l + r
    return jac_row<MODE>(z, t0, t1, t2, t3, up, dn, a, c, E.zh, E.zl);
}

template <int R, int NW, int MODE, bool EDGE>
__device__ __forceinline__ void jacobi_tile_body(const TileArgs& A, int kx, int ky, float4 (&v)[R + 1],
                                                 float4 (&z)[R], float4* __restrict__ halo) {
    constexpr int TH = R * NW;
    constexpr int HB = (NW * 2 + 2) * 32;  // float4 slots per halo buffer (one guard row at each end)
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const Geom& g = A.g;
    const int ox = kx * A.SX, oy = ky * A.SY;
    const int gx0 = ox + 4 * lane;  // global column of component .x
    const int gyw = oy + wi * R;    // LOCAL row of this warp's row 0
    const int gyg = gyw + g.gy0;    // its GLOBAL row (boundary logic is global, addressing local)
    const bool negx = (A.b == B_NEGX), negy = (A.b == B_NEGY);
    const float a = A.a, c = A.c;

    EdgeFlags E;
    E.l1 = E.r0 = E.r1 = E.r2 = E.r3 = false;
    E.negx = negx; E.negy = negy; E.h = g.gh;
    E.zh = A.zh; E.zl = A.zl;
    if (EDGE) {
        E.l1 = (gx0 + 1 == 1);  // ox == 0 && lane == 0 (tile origins are multiples of 4)
        E.r0 = (gx0 + 0 == g.w - 2); E.r1 = (gx0 + 1 == g.w - 2); E.r2 = (gx0 + 2 == g.w - 2); E.r3 = (gx0 + 3 == g.w - 2);
    }

    if (A.fuse_src) {  // add_source: x0[k] = x0[k] + s[k]*dt on every cell (border included)
        const float dt = A.src_dt;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            z[r].x = fadd(z[r].x, fmul(v[r + 1].x, dt));
            z[r].y = fadd(z[r].y, fmul(v[r + 1].y, dt));
            z[r].z = fadd(z[r].z, fmul(v[r + 1].z, dt));
            z[r].w = fadd(z[r].w, fmul(v[r + 1].w, dt));
        }
    }

    // My top-row slot in halo buffer 0 (+32: my bottom row).  Slot -32 is the bottom row of the
    // warp above, slot +64 the top row of the warp below; warp 0 / NW-1 hit the guard rows
    // (stale values: those rows are outside the tile and never feed a stored cell).
    float4* const hw = halo + (wi * 2 + 1) * 32 + lane;

    for (int s = 1; s <= A.T; s += 2) {
        {   // ---- down sweep: rows in v[1..R] -> v[0..R-1] ----
            hw[HB] = v[1];
            hw[HB + 32] = v[R];
            __syncthreads();
            v[0] = hw[HB - 32];
            const float4 hbot = hw[HB + 64];
            const bool sub = EDGE && (s > 1);
#pragma unroll
            for (int r = 1; r <= R; ++r)
                v[r - 1] = jac_row_full<MODE, EDGE>(v[r], v[r - 1], (r == R) ? hbot : v[r + 1], z[r - 1], a, c, E,
                                                    sub, gyg + r - 1);
        }
        if (s + 1 > A.T) {  // odd T: move the rows back to v[1..R] for the store phase
#pragma unroll
            for (int r = R; r >= 1; --r) v[r] = v[r - 1];
            break;
        }
        {   // ---- up sweep: rows in v[0..R-1] -> v[1..R] ----
            hw[0] = v[0];
            hw[32] = v[R - 1];
            __syncthreads();
            v[R] = hw[64];
            const float4 htop = hw[-32];
#pragma unroll
            for (int r = R - 1; r >= 0; --r)
                v[r + 1] = jac_row_full<MODE, EDGE>(v[r], (r == 0) ? htop : v[r - 1], v[r + 1], z[r], a, c, E, EDGE,
                                                    gyg + r);
        }
    }
    float4* const xr = v + 1;  // rows 0..R-1 of this warp

    // ---- store the cells that are H away from every non-domain tile edge -----------------
    (void)FULL;
    const bool lastx = (kx == A.ntx - 1);
    const int xlo = (kx == 0) ? 0 : ox + A.H, xhi = lastx ? g.w : ox + 128 - A.H;   // [xlo, xhi)
    // Rows (LOCAL indices): a local edge that is the domain edge needs no halo; one that is a slab
    // cut is stale H rows deep.  Only interior rows are stored here; border rows follow below.
    const bool top_dom = (g.gy0 == 0), bot_dom = (g.gy0 + g.h == g.gh);
    const int ylo = (ky == 0 && top_dom) ? 1 : oy + A.H;
    const int yhi = (ky == A.nty - 1) ? (bot_dom ? g.h - 1 : min(oy + TH - A.H, g.h - A.H)) : oy + TH - A.H;
    const bool v0 = gx0 + 0 >= xlo && gx0 + 0 < xhi, v1 = gx0 + 1 >= xlo && gx0 + 1 < xhi;
    const bool v2 = gx0 + 2 >= xlo && gx0 + 2 < xhi, v3 = gx0 + 3 >= xlo && gx0 + 3 < xhi;
    const bool vall = v0 && v1 && v2 && v3;
    if (!EDGE) {
        // Interior tile: [xlo, xhi) = [ox+H, ox+128-H) with H even, so a lane stores its whole
        // float4, its low or high half (8-byte aligned), or nothing.  One branch on the lane's
        // mode outside the row loop; inside it only a predicated store per row.
        const int lo = A.H - 4 * lane, hi = 128 - A.H - 4 * lane;  // my components [lo, hi) are stored
        const int mode = (lo <= 0 && hi >= 4) ? 1 : (lo <= 0 && hi == 2) ? 2 : (lo == 2 && hi >= 4) ? 3 : 0;
        const unsigned rlo = (unsigned)max(0, ylo - gyw), rcnt = (unsigned)max(0, min(R, yhi - gyw) - (int)rlo);
        float* const o0 = A.out + (size_t)gyw * g.pitch + gx0;
        float* const z0 = (A.fuse_src && A.z_out) ? A.z_out + (size_t)gyw * g.pitch + gx0 : nullptr;
        const size_t pitch = g.pitch;
        if (mode == 1) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if ((unsigned)r - rlo < rcnt) *reinterpret_cast<float4*>(o0 + r * pitch) = xr[r];
            if (z0) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if ((unsigned)r - rlo < rcnt) *reinterpret_cast<float4*>(z0 + r * pitch) = z[r];
            }
        } else if (mode != 0) {
            const int sh = (mode == 3) ? 2 : 0;  // which half of the float4
#pragma unroll
            for (int r = 0; r < R; ++r)
                if ((unsigned)r - rlo < rcnt)
                    *reinterpret_cast<float2*>(o0 + r * pitch + sh) = sh ? make_float2(xr[r].z, xr[r].w) : make_float2(xr[r].x, xr[r].y);
            if (z0) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if ((unsigned)r - rlo < rcnt)
                        *reinterpret_cast<float2*>(z0 + r * pitch + sh) = sh ? make_float2(z[r].z, z[r].w) : make_float2(z[r].x, z[r].y);
            }
        }
        return;
    }
    if (A.fuse_src && A.z_out) {  // the updated right-hand side, all rows of my output band (border rows too)
        const int zlo = (ky == 0 && top_dom) ? 0 : ylo, zhi = (yhi == g.h - 1 && bot_dom) ? g.h : yhi;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int gy = gyw + r;
            if (gy < zlo || gy >= zhi) continue;
            float* row = A.z_out + (size_t)gy * g.pitch;
            if (vall) {
                *reinterpret_cast<float4*>(row + gx0) = z[r];
            } else {
                if (v0) row[gx0 + 0] = z[r].x;
                if (v1) row[gx0 + 1] = z[r].y;
                if (v2) row[gx0 + 2] = z[r].z;
                if (v3) row[gx0 + 3] = z[r].w;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int gy = gyw + r;
        if (gy < ylo || gy >= yhi) continue;  // warp-uniform
        float4 f = xr[r];
        // column borders of this interior row (lib.rs:31-32); vleft = value at gx0-1
        const float vleft = __shfl_up_sync(FULL, f.w, 1);
        if (gx0 == 0) f.x = sgn(negx, f.y);
        if (gx0 + 0 == g.w - 1) f.x = sgn(negx, vleft);
        if (gx0 + 1 == g.w - 1) f.y = sgn(negx, f.x);
        if (gx0 + 2 == g.w - 1) f.z = sgn(negx, f.y);
        if (gx0 + 3 == g.w - 1) f.w = sgn(negx, f.z);
        float* row = A.out + (size_t)gy * g.pitch;
        if (vall) {
            *reinterpret_cast<float4*>(row + gx0) = f;
        } else {
            if (v0) row[gx0 + 0] = f.x;
            if (v1) row[gx0 + 1] = f.y;
            if (v2) row[gx0 + 2] = f.z;
            if (v3) row[gx0 + 3] = f.w;
        }
        const int gyglob = gy + g.gy0;
        if (gyglob == 1 || gyglob == g.gh - 2) {
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
                if (side == 0 && gyglob != 1) continue;
                if (side == 1 && gyglob != g.gh - 2) continue;
                float* brow = A.out + (size_t)(side == 0 ? gy - 1 : gy + 1) * g.pitch;
                if (v0) brow[gx0 + 0] = e.x;
                if (v1) brow[gx0 + 1] = e.y;
                if (v2) brow[gx0 + 2] = e.z;
                if (v3) brow[gx0 + 3] = e.w;
            }
        }
    }
}

template <int R, int NW>
constexpr size_t jacobi_tile_smem_bytes() {
    // TMA staging of the next tile's x0 and x + halo exchange (2 buffers, NW*2+2 rows of 512 B) + mbarrier
    return (size_t)2 * R * NW * 512 + (size_t)2 * (NW * 2 + 2) * 512 + 32;
}

template <int R, int NW, int MODE>
__global__ void __launch_bounds__(NW * 32, 1) k_jacobi_tile(const __grid_constant__ TileArgs A) {
    pdl_enter();
    constexpr int TH = R * NW;
    constexpr uint32_t TILE_BYTES = TH * 512;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* const zs = reinterpret_cast<float*>(smem_raw);                        // [TH][128] x0 staging (TMA dst)
    float* const xs = reinterpret_cast<float*>(smem_raw + (size_t)TILE_BYTES);   // [TH][128] x staging  (TMA dst)
    float4* const halo = reinterpret_cast<float4*>(smem_raw + (size_t)2 * TILE_BYTES);
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)2 * TILE_BYTES + (size_t)2 * (NW * 2 + 2) * 512);
    volatile int* const s_item = reinterpret_cast<volatile int*>(mbar + 1);       // [2]: item of this / the next tile
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int ntiles = A.ntx * A.nty;
    const uint32_t tx_bytes = A.x_is_zero ? TILE_BYTES : 2 * TILE_BYTES;

    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        const int item = atomicAdd(&A.sched[0], 1);
        s_item[0] = item;
        if (item < ntiles) {
            int kx, ky;
            tile_of_item(item, A.ntx, A.nty, kx, ky);
            mbar_expect_tx(mbar, tx_bytes);
            tma_load_2d(zs, &A.tm_x0, kx * A.SX, ky * A.SY, mbar);
            if (!A.x_is_zero) tma_load_2d(xs, &A.tm_x, kx * A.SX, ky * A.SY, mbar);
        }
    }
    __syncthreads();  // barrier init + first work item visible to every thread

    uint32_t phase = 0;
    int slot = 0;
    for (int item = s_item[0]; item < ntiles; item = s_item[slot]) {
        int kx, ky;
        tile_of_item(item, A.ntx, A.nty, kx, ky);
        // Claim the next work item now: the atomic's round trip hides behind the wait + copy below.
        int next = 0;
        if (threadIdx.x == 0) next = atomicAdd(&A.sched[0], 1);
        mbar_wait(mbar, phase);  // this tile's x0 (and x) have landed in the staging buffers
        phase ^= 1u;

        // Staging -> registers: the iterate and the right-hand side stay there for all T sweeps.
        float4 v[R + 1], z[R];
        {
            const float4* xrow = reinterpret_cast<const float4*>(xs) + (wi * R) * 32 + lane;
            const float4* zrow = reinterpret_cast<const float4*>(zs) + (wi * R) * 32 + lane;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                v[r + 1] = A.x_is_zero ? make_float4(0.f, 0.f, 0.f, 0.f) : xrow[r * 32];
                z[r] = zrow[r * 32];
            }
            v[0] = v[1];
        }
        __syncthreads();  // staging buffers are free again (and the previous tile's halo reads are done)
        slot ^= 1;
        if (threadIdx.x == 0) {  // prefetch the next tile behind this tile's sweeps
            s_item[slot] = next;  // read by everyone after the sweeps' barriers
            if (next < ntiles) {
                int nkx, nky;
                tile_of_item(next, A.ntx, A.nty, nkx, nky);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before async writes
                mbar_expect_tx(mbar, tx_bytes);
                tma_load_2d(zs, &A.tm_x0, nkx * A.SX, nky * A.SY, mbar);
                if (!A.x_is_zero) tma_load_2d(xs, &A.tm_x, nkx * A.SX, nky * A.SY, mbar);
            }
        }
        const int ox = kx * A.SX, oy = ky * A.SY;
        const bool edge = (kx == 0) || (ox + 128 >= A.g.w - 1) || (oy + A.g.gy0 <= 1) || (oy + TH + A.g.gy0 >= A.g.gh - 1) ||
                          (ky == 0) || (ky == A.nty - 1);  // slab cuts take the general store path too
        if (edge)
            jacobi_tile_body<R, NW, MODE, true>(A, kx, ky, v, z, halo);
        else
            jacobi_tile_body<R, NW, MODE, false>(A, kx, ky, v, z, halo);
    }
    // The last CTA to leave resets the scheduler for the next launch on this stream.
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&A.sched[1], 1) == (int)gridDim.x - 1) {
            A.sched[0] = 0;
            A.sched[1] = 0;
            __threadfence();
        }
    }
}

// Exhaustive check of the MODE_RECIP division for one divisor: all 2^23 significands, two binades, both signs.
__global__ void k_recip_verify(float c, float zh, float zl, int variant, int* bad) {
    const unsigned m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (1u << 23)) return;
    int b = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned bits = ((k & 1) ? 0x42000000u : 0x3f800000u) | ((k & 2) ? 0x80000000u : 0u) | m;
        const float v = __uint_as_float(bits);
        float q;
        if (variant == MODE_RECIP) {
            q = __fmaf_rn(v, zh, __fmul_rn(v, zl));
        } else {  // MODE_RECIP3
            const float q0 = __fmul_rn(v, zh);
            q = __fmaf_rn(__fmaf_rn(q0, -c, v), zh, q0);
        }
        b |= (__float_as_uint(q) != __float_as_uint(__fdiv_rn(v, c)));
    }
    if (b) atomicOr(bad, 1);
}

// ---- host side -----------------------------------------------------------------------------
// c -> (usable, zh, zl); filled by recip_for() the first time a divisor is seen (synchronises once).
struct RecipInfo { bool ok; int mode; float zh, zl; };
static std::mutex g_recip_mu;
static std::unordered_map<uint32_t, RecipInfo> g_recip;
static bool recip_for(float c, RecipInfo* out) {
    uint32_t key;
    memcpy(&key, &c, 4);
    std::lock_guard<std::mutex> lk(g_recip_mu);
    auto it = g_recip.find(key);
    if (it != g_recip.end()) { *out = it->second; return it->second.ok; }
    RecipInfo r{false, MODE_GENERAL, 0.f, 0.f};
    int e = 0;
    (void)frexpf(c, &e);
    if (c == c && c != 0.f && e > -30 && e < 30) {  // finite, normal, moderate magnitude
        r.zh = 1.0f / c;
        r.zl = (float)(1.0 / (double)c - (double)r.zh);
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        int* bad = nullptr;
        // never synchronise inside a stream capture: leave the divisor unverified (general path) for now
        if (cudaStreamIsCapturing(cudaStreamPerThread, &cs) == cudaSuccess && cudaMalloc(&bad, sizeof(int)) == cudaSuccess) {
            for (int variant : {(int)MODE_RECIP, (int)MODE_RECIP3}) {
                int h = 1;
                if (cudaMemset(bad, 0, sizeof(int)) != cudaSuccess) break;
                k_recip_verify<<<(1u << 23) / 256, 256>>>(c, r.zh, r.zl, variant, bad);
                if (cudaMemcpy(&h, bad, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) break;
                if (h == 0) {
                    r.ok = true; r.mode = variant;
                    break;
                }
            }
            cudaFree(bad);
        }
    }
    g_recip.emplace(key, r);
    *out = r;
    return r.ok;
}

struct TileCfg { int R, NW; };
static TileCfg g_tile_cfg = {8, 16};
static int g_tile_default_T = 10;

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

struct TmKey {
    const void* p; size_t pitch; int h; int th;
    bool operator==(const TmKey& o) const { return p == o.p && pitch == o.pitch && h == o.h && th == o.th; }
};
struct TmKeyHash {
    size_t operator()(const TmKey& k) const {
        return std::hash<const void*>()(k.p) ^ (k.pitch * 1315423911u) ^ ((size_t)k.h << 20) ^ (size_t)k.th;
    }
};
static std::mutex g_tm_mu;
static std::unordered_map<TmKey, CUtensorMap, TmKeyHash> g_tm_cache;

// Tensor map of a field for (128 x th) boxes.  Returns 0 or -4 (CUDA error).
static int tile_tensor_map(const float* base, size_t pitch, int h, int th, CUtensorMap* out) {
    std::lock_guard<std::mutex> lk(g_tm_mu);
    const TmKey key{base, pitch, h, th};
    auto it = g_tm_cache.find(key);
    if (it != g_tm_cache.end()) { *out = it->second; return 0; }
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return -4;
    const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)h};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
    const cuuint32_t box[2] = {128u, (cuuint32_t)th};
    const cuuint32_t estr[2] = {1u, 1u};
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -4;
    if (g_tm_cache.size() > 4096) g_tm_cache.clear();
    g_tm_cache.emplace(key, tm);
    *out = tm;
    return 0;
}

// Per-stream scheduler words (zeroed once; every launch leaves them zero again).
static std::mutex g_sched_mu;
static std::unordered_map<cudaStream_t, int*> g_sched;
static int* tile_sched_for(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_sched_mu);
    auto it = g_sched.find(st);
    if (it != g_sched.end()) return it->second;
    int* p = nullptr;
    if (cudaMalloc(&p, 2 * sizeof(int)) != cudaSuccess) return nullptr;
    // Zeroed ON THE STREAM that will use them: cudaMemset would run on the legacy default stream, which does not
    // order with non-blocking streams -- the first kernel could read the words before they are cleared.
    if (cudaMemsetAsync(p, 0, 2 * sizeof(int), st) != cudaSuccess) { cudaFree(p); return nullptr; }
    g_sched.emplace(st, p);
    return p;
}

static inline int tiles_needed(int extent, int tile, int stride) {
    if (extent <= tile) return 1;
    return (extent - tile + stride - 1) / stride + 1;
}

static void tile_sched_release(cudaStream_t st) {  // the stream is going away (handle destroyed)
    std::lock_guard<std::mutex> lk(g_sched_mu);
    auto it = g_sched.find(st);
    if (it == g_sched.end()) return;
    cudaFree(it->second);
    g_sched.erase(it);
}

static inline int jacobi_tile_max_T(int R, int NW) {
    const int th = R * NW;
    int H = (th < 128 ? th : 128) / 2 - 2;  // keep at least 4 output rows/cols per tile
    return H & ~1;
}

static inline long jacobi_tile_count(const Geom& g, int R, int NW, int T) {
    const int H = T + (T & 1), TH = R * NW;
    return (long)tiles_needed(g.w, 128, 128 - 2 * H) * tiles_needed(g.h, TH, TH - 2 * H);
}

// Tile shape and fused depth per launch.  The kernel's cost per tile and sweep is proportional to the
// tile's rows (it is issue bound), so as long as a tiling leaves SMs idle a SMALLER tile finishes sooner.
// Measured scene step (tools/small_grid_cfg_probe.py, us; 128x128 tiles at T = 20 in brackets):
//   250^2: 86 with 128x32 tiles (162)    512^2: 104 with 128x48 (167)
//   768^2: 119 with 128x64 (169)         1024^2: 156 with 128x96 (175)       all at T = 10.
// From ~1200^2 upwards every SM has a tile anyway and the least halo recompute wins: 128x128, T = 10.
// Automatic choice (until fruid_set_tile_config pins a shape; always the pinned / default shape in slab mode,
// whose halo bookkeeping assumes it): the first of these that gives every tile its own SM.
static bool g_tile_auto = true;
struct TilePlan { int R, NW, T; };
static const TilePlan kSmallGridPlans[] = {{2, 16, 10}, {3, 16, 10}, {2, 32, 10}, {4, 24, 10}, {8, 16, 20}};

static inline int jacobi_tile_default_fused(const Geom& g, int iters) {
    int T = g_tile_default_T;
    const bool whole = (g.gh == g.h);
    if (whole && g_tile_auto) {
        for (const TilePlan& p : kSmallGridPlans) {
            const int t = p.T > iters ? iters : p.T;
            if (jacobi_tile_count(g, p.R, p.NW, t) <= 148) return t < 1 ? 1 : t;
        }
    } else if (whole) {
        // pinned shape: when the tiling at the default depth cannot even give every SM one tile, fuse a
        // whole 20-sweep solve per launch (more halo recompute on SMs that would otherwise idle, half the launches)
        if (jacobi_tile_count(g, g_tile_cfg.R, g_tile_cfg.NW, T) < 148) T = 20;
    }
    const int mx = jacobi_tile_max_T(g_tile_cfg.R, g_tile_cfg.NW);
    if (T > mx) T = mx;
    if (T > iters) T = iters;
    return T < 1 ? 1 : T;
}

// Shape for a launch of T fused sweeps.
static inline TileCfg jacobi_tile_shape(const Geom& g, int T) {
    if (g_tile_auto && g.gh == g.h)
        for (const TilePlan& p : kSmallGridPlans)
            if (T <= jacobi_tile_max_T(p.R, p.NW) && jacobi_tile_count(g, p.R, p.NW, T) <= 148) return TileCfg{p.R, p.NW};
    return g_tile_cfg;
}

template <int R, int NW, int MODE>
static int jacobi_tile_launch_mode(const TileArgs& A, cudaStream_t st) {
    constexpr size_t smem = jacobi_tile_smem_bytes<R, NW>();
    static bool attr_set = false;  // per instantiation
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_jacobi_tile<R, NW, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -4;
        attr_set = true;
    }
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            return -4;
    }
    const int ntiles = A.ntx * A.nty;
    dim3 grid(ntiles < num_sms ? ntiles : num_sms), block(NW * 32);
    return launch_pdl(true, k_jacobi_tile<R, NW, MODE>, grid, block, smem, st, A) == cudaSuccess ? 0 : -4;
}

template <int R, int NW>
static int jacobi_tile_launch_cfg(const TileArgs& A, int mode, cudaStream_t st) {
    if (mode == MODE_GENERAL) return jacobi_tile_launch_mode<R, NW, MODE_GENERAL>(A, st);
    if (mode == MODE_POW2) return jacobi_tile_launch_mode<R, NW, MODE_POW2>(A, st);
    if (mode == MODE_POISSON) return jacobi_tile_launch_mode<R, NW, MODE_POISSON>(A, st);
    if (mode == MODE_RECIP) return jacobi_tile_launch_mode<R, NW, MODE_RECIP>(A, st);
    if (mode == MODE_RECIP3) return jacobi_tile_launch_mode<R, NW, MODE_RECIP3>(A, st);
    return jacobi_tile_launch_mode<R, NW, MODE_ONE>(A, st);
}

#define FRUID_TILE_CONFIGS(X) X(8, 16) X(10, 16) X(6, 16) X(6, 20) X(7, 20) X(5, 24) X(4, 24) X(8, 12) X(4, 16) X(2, 32) X(2, 16) X(4, 8) X(3, 16)

// Returns 0 or a negative fruid_status (-7 = bad argument, -4 = CUDA error).
static int jacobi_tile_launch(float* out, const float* x, const float* x0, const Geom& g, float a, float c, int b, int T,
                              int x_is_zero, cudaStream_t st, int fuse_src = 0, float src_dt = 0.f, float* z_out = nullptr) {
    if (T < 1) return -7;
    const TileCfg cfg = jacobi_tile_shape(g, T);
    const int R = cfg.R, NW = cfg.NW, TH = R * NW;
    if (T > jacobi_tile_max_T(R, NW)) return -7;
    TileArgs A;
    if (int r = tile_tensor_map(x0, g.pitch, g.h, TH, &A.tm_x0)) return r;
    if (int r = tile_tensor_map(x, g.pitch, g.h, TH, &A.tm_x)) return r;
    A.out = out; A.x = x; A.g = g; A.a = a; A.b = b; A.T = T; A.x_is_zero = x_is_zero;
    A.zh = A.zl = 0.f;
    A.fuse_src = fuse_src; A.src_dt = src_dt; A.z_out = z_out;
    A.sched = tile_sched_for(st);
    if (!A.sched) return -4;
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
        if (m == 0.5f && e > -100 && e < 100) {  // c = 2^k: multiply by the exact reciprocal
            mode = (a == 1.0f) ? MODE_POISSON : MODE_POW2;
            A.c = 1.0f / c;
        } else {
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            RecipInfo ri;
            const bool capturing = cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone;
            bool known;
            {
                uint32_t key; memcpy(&key, &c, 4);
                std::lock_guard<std::mutex> lk(g_recip_mu);
                known = g_recip.count(key) != 0;
            }
            if ((known || !capturing) && recip_for(c, &ri)) { mode = ri.mode; A.zh = ri.zh; A.zl = ri.zl; }
        }
    }
#define FRUID_TILE_CASE(r, nw) if (R == r && NW == nw) return jacobi_tile_launch_cfg<r, nw>(A, mode, st);
    FRUID_TILE_CONFIGS(FRUID_TILE_CASE)
#undef FRUID_TILE_CASE
    return -7;
}

static inline int jacobi_tile_set_cfg(int R, int NW, int T) {
    if (R == 0 && NW == 0) {  // back to the automatic choice
        g_tile_auto = true; g_tile_cfg = {8, 16}; g_tile_default_T = T > 0 ? T : 10;
        return 0;
    }
#define FRUID_TILE_OK(r, nw) if (R == r && NW == nw) { g_tile_cfg = {R, NW}; g_tile_auto = false; if (T > 0) g_tile_default_T = T; return 0; }
    FRUID_TILE_CONFIGS(FRUID_TILE_OK)
#undef FRUID_TILE_OK
    return -7;
}

}  // namespace fruid
