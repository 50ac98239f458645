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
// MODE_ZEROA is c == 1 with a == +-0 (diffusion with diff = 0, the reference scene: main.rs:110-118): the product
// a*sum is then EXACT (+-0, or NaN for an infinite or NaN sum), so fma(a, sum, x0) rounds the same real number as the
// unfused x0 + a*sum and obeys the same signed-zero rules -- one packed FFMA2 instead of two multiplies and an add.
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

#include "common.cuh"

namespace fruid {

enum : int { MODE_GENERAL = 0, MODE_POW2 = 1, MODE_ONE = 2, MODE_POISSON = 3, MODE_RECIP = 4, MODE_RECIP3 = 5, MODE_ZEROA = 6 };

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
    int pdl_trigger;  // programmatic dependent launch: let the successor be scheduled early (host switch)
    int dbg_delay_ns; // FRUID_DEBUG builds: the lowest warps stall this long before storing (0 = off)
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
    if (MODE == MODE_ZEROA) {  // x0 + a*sum with an exact product: see the header
        s01 = __ffma2_rn(make_float2(a, a), s01, make_float2(z.x, z.y));
        s23 = __ffma2_rn(make_float2(a, a), s23, make_float2(z.z, z.w));
        return make_float4(s01.x, s01.y, s23.x, s23.y);
    }
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

// Boundary handling inside the sweeps ("set_bnd in registers", lib.rs:27-44 after every sweep, lib.rs:69).
// A tile that contains a domain border keeps the border cells of its register copy equal to what set_bnd would
// have written: after every sweep the thread (column borders) or warp (row borders) that holds a border cell
// overwrites it with sigma * (adjacent interior cell of the new iterate).  Sweep 1 therefore reads the incoming
// border as loaded (lib.rs:56-59 read x before any set_bnd) and later sweeps read what set_bnd left there; the
// stencil itself is the interior code, unchanged.  Corners are never read by the 5-point stencil; the store phase
// produces them (and the border rows / columns of the final iterate) exactly.
//   EDGE = 0  interior tile: nothing to do
//   EDGE = 1  the tile holds the top and/or bottom domain row (and no left/right domain column): a few
//             warp-uniform fixups per SWEEP
//   EDGE = 2  the tile also holds the left and/or right domain column: per-lane selects per row as well
struct EdgeFlags {
    bool negx, negy;
    bool pl;            // my .x is border column 0
    bool p1, p2, p3;    // my .y / .z / .w is border column w-1 (its source, column w-2, is my previous component)
    bool p0r;           // my .w is column w-2 and border column w-1 is the next lane's .x: mirror on read
    bool top;           // this warp holds global rows 0 (border) and 1 in its first two rows
    unsigned bot;       // bit r set: my row r is global row gh-2, whose lower neighbour is border row gh-1
    float zh, zl;
};
// sigma * v selected without touching the FP32 pipe (sign flip = integer XOR; the select is an ALU-pipe FSEL).
__device__ __forceinline__ float sel_mirror(bool take, unsigned flip, float src, float keep) {
    return take ? __uint_as_float(__float_as_uint(src) ^ flip) : keep;
}
__device__ __forceinline__ float4 sel_mirror4(bool take, unsigned flip, float4 src, float4 keep) {
    return make_float4(sel_mirror(take, flip, src.x, keep.x), sel_mirror(take, flip, src.y, keep.y),
                       sel_mirror(take, flip, src.z, keep.z), sel_mirror(take, flip, src.w, keep.w));
}

// New value of one float4 row from its old value cc, the rows above/below and x0 (z).  Boundary rows and columns
// are handled by selects on the operands / results (no branches: a divergent branch per row costs more than the
// whole stencil):
//   rows     TOPROW (a constant after unrolling: this is warp-row 1): if the warp holds the top border, `up` = sy * cc (lib.rs:36);
//            any row: if it is global row gh-2 (E.bot), `dn` = sy * cc (lib.rs:37)
//   columns  the border cells of the NEW row are set to sx * their neighbour (lib.rs:31-32), so the next sweep reads
//            them like any other cell; a border column in the next lane's .x is mirrored on read instead (p0r)
// `sub` is false in the first sweep of a launch, which reads the incoming border as loaded (lib.rs:56-59).
template <int MODE, int EDGE>
__device__ __forceinline__ float4 jac_row_full(float4 cc, float4 up, float4 dn, const float4 z, float a, float c,
                                               const EdgeFlags& E, bool sub, bool isbot, const bool TOPROW) {
    const unsigned FULL = 0xffffffffu;
    float lft = __shfl_up_sync(FULL, cc.w, 1);
    float rgt = __shfl_down_sync(FULL, cc.x, 1);
    if (EDGE >= 1) {
        const unsigned fy = E.negy ? 0x80000000u : 0u;
        if (TOPROW) up = sel_mirror4(sub && E.top, fy, cc, up);
        dn = sel_mirror4(sub && isbot, fy, cc, dn);
    }
    const unsigned fx = E.negx ? 0x80000000u : 0u;
    if (EDGE == 2) rgt = sel_mirror(sub && E.p0r, fx, cc.w, rgt);
    const float t0 = fadd(lft, cc.y), t1 = fadd(cc.x, cc.z), t2 = fadd(cc.y, cc.w), t3 = fadd(cc.z, rgt);  // l + r
    float4 nv = jac_row<MODE>(z, t0, t1, t2, t3, up, dn, a, c, E.zh, E.zl);
    if (EDGE == 2) {  // column borders of the new iterate
        nv.x = sel_mirror(E.pl, fx, nv.y, nv.x);
        nv.y = sel_mirror(E.p1, fx, nv.x, nv.y);
        nv.z = sel_mirror(E.p2, fx, nv.y, nv.z);
        nv.w = sel_mirror(E.p3, fx, nv.z, nv.w);
    }
    return nv;
}

template <int R, int NW, int MODE, int EDGE>
__device__ __forceinline__ void jacobi_tile_body(const TileArgs& A, int kx, int ky, float4 (&v)[R + 1],
                                                 float4 (&z)[R], float4* __restrict__ halo) {
    static_assert(R >= 2, "a warp must hold a border row together with its neighbour");
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
    E.negx = negx; E.negy = negy;
    E.pl = E.p1 = E.p2 = E.p3 = E.p0r = E.top = false;
    E.bot = 0u;
    E.zh = A.zh; E.zl = A.zl;
    if (EDGE >= 1) {
        E.top = (gyg == 0);
        const int rb = g.gh - 2 - gyg;  // my row that is global row gh-2, if any
        if (rb >= 0 && rb < R) E.bot = 1u << rb;
    }
    if (EDGE == 2) {
        E.pl = (gx0 == 0);
        E.p1 = (gx0 + 1 == g.w - 1); E.p2 = (gx0 + 2 == g.w - 1); E.p3 = (gx0 + 3 == g.w - 1);
        E.p0r = (gx0 + 4 == g.w - 1);
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
            for (int r = 1; r <= R; ++r)  // warp-row r-1
                v[r - 1] = jac_row_full<MODE, EDGE>(v[r], v[r - 1], (r == R) ? hbot : v[r + 1], z[r - 1], a, c, E, sub,
                                                    EDGE && ((E.bot >> (r - 1)) & 1u), r == 2);
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
            for (int r = R - 1; r >= 0; --r)  // warp-row r
                v[r + 1] = jac_row_full<MODE, EDGE>(v[r], (r == 0) ? htop : v[r - 1], v[r + 1], z[r], a, c, E, EDGE != 0,
                                                    EDGE && ((E.bot >> r) & 1u), r == 1);
        }
    }
    float4* const xr = v + 1;  // rows 0..R-1 of this warp
    dbg_tail_delay_ns(wi, A.dbg_delay_ns);

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
        // Interior tile: no tile edge is a domain edge, so the stored cells are the tile-local box [H, 128-H) x [H, TH-H)
        // whatever the tile.  H is even: a lane stores its low half, its high half, both or nothing -- two predicated
        // 8-byte stores per row, the same instruction stream in every lane.
        const int c0 = A.H - 4 * lane, c1 = 128 - A.H - 4 * lane;  // my components [c0, c1) are stored
        const bool slo = (c0 <= 0) && (c1 >= 2), shi = (c0 <= 2) && (c1 >= 4);
        const int r0 = A.H - wi * R, r1 = TH - A.H - wi * R;        // my rows [r0, r1) are stored
        const size_t pitch = g.pitch;
        float* o = A.out + (size_t)gyw * pitch + gx0;
        float* zo = (A.fuse_src && A.z_out) ? A.z_out + (size_t)gyw * pitch + gx0 : nullptr;
        if (r0 <= 0 && r1 >= R) {  // all my rows (warp-uniform; every warp but the first and last two of a tile)
            if (slo && shi) {      // 26 of 32 lanes: whole float4s
#pragma unroll
                for (int r = 0; r < R; ++r, o += pitch) *reinterpret_cast<float4*>(o) = xr[r];
                if (zo) {
#pragma unroll
                    for (int r = 0; r < R; ++r, zo += pitch) *reinterpret_cast<float4*>(zo) = z[r];
                }
            } else if (slo || shi) {
                const int sh = shi ? 2 : 0;
#pragma unroll
                for (int r = 0; r < R; ++r, o += pitch)
                    *reinterpret_cast<float2*>(o + sh) = shi ? make_float2(xr[r].z, xr[r].w) : make_float2(xr[r].x, xr[r].y);
                if (zo) {
#pragma unroll
                    for (int r = 0; r < R; ++r, zo += pitch)
                        *reinterpret_cast<float2*>(zo + sh) = shi ? make_float2(z[r].z, z[r].w) : make_float2(z[r].x, z[r].y);
                }
            }
        } else if (r1 > 0 && r0 < R) {
#pragma unroll
            for (int r = 0; r < R; ++r, o += pitch) {
                if (r < r0 || r >= r1) continue;  // warp-uniform
                if (slo) *reinterpret_cast<float2*>(o) = make_float2(xr[r].x, xr[r].y);
                if (shi) *reinterpret_cast<float2*>(o + 2) = make_float2(xr[r].z, xr[r].w);
            }
            if (zo) {
#pragma unroll
                for (int r = 0; r < R; ++r, zo += pitch) {
                    if (r < r0 || r >= r1) continue;
                    if (slo) *reinterpret_cast<float2*>(zo) = make_float2(z[r].x, z[r].y);
                    if (shi) *reinterpret_cast<float2*>(zo + 2) = make_float2(z[r].z, z[r].w);
                }
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

// One tile: staging -> registers, prefetch of the next tile, T sweeps, stores.
template <int R, int NW, int MODE, int EDGE>
__device__ __forceinline__ void jacobi_tile_run(const TileArgs& A, int kx, int ky, float* zs, float* xs, float4* halo,
                                                uint64_t* mbar, volatile int* s_next, int next, int ntiles, uint32_t tx_bytes) {
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
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
    if (threadIdx.x == 0) {  // prefetch the next tile behind this tile's sweeps
        *s_next = next;      // read by everyone after the sweeps' barriers
        if (next < ntiles) {
            int nkx, nky;
            tile_of_item(next, A.ntx, A.nty, nkx, nky);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before async writes
            mbar_expect_tx(mbar, tx_bytes);
            tma_load_2d(zs, &A.tm_x0, nkx * A.SX, nky * A.SY, mbar);
            if (!A.x_is_zero) tma_load_2d(xs, &A.tm_x, nkx * A.SX, nky * A.SY, mbar);
        }
    }
    jacobi_tile_body<R, NW, MODE, EDGE>(A, kx, ky, v, z, halo);
}

template <int R, int NW>
constexpr size_t jacobi_tile_smem_bytes() {
    // TMA staging of the next tile's x0 and x + halo exchange (2 buffers, NW*2+2 rows of 512 B) + mbarrier
    return (size_t)2 * R * NW * 512 + (size_t)2 * (NW * 2 + 2) * 512 + 32;
}

template <int R, int NW, int MODE>
__global__ void __launch_bounds__(NW * 32, 1) k_jacobi_tile(const __grid_constant__ TileArgs A) {
    pdl_enter_arg(A.pdl_trigger);
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

        // The three boundary classes (see jacobi_tile_body) are three self-contained code paths, each loading its own
        // registers from the staging buffers: sharing the register copy across them costs register moves in all three.
        const int ox = kx * A.SX, oy = ky * A.SY;
        const bool xedge = (kx == 0) || (ox + 128 >= A.g.w - 1);
        const bool yedge = (oy + A.g.gy0 <= 1) || (oy + TH + A.g.gy0 >= A.g.gh - 1) || (ky == 0) || (ky == A.nty - 1);
        slot ^= 1;
        if (xedge)
            jacobi_tile_run<R, NW, MODE, 2>(A, kx, ky, zs, xs, halo, mbar, s_item + slot, next, ntiles, tx_bytes);
        else if (yedge)
            jacobi_tile_run<R, NW, MODE, 1>(A, kx, ky, zs, xs, halo, mbar, s_item + slot, next, ntiles, tx_bytes);
        else
            jacobi_tile_run<R, NW, MODE, 0>(A, kx, ky, zs, xs, halo, mbar, s_item + slot, next, ntiles, tx_bytes);
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


// Launchers of the instantiations, one translation unit per group of tile shapes (tile_inst_*.cu) so that the
// library builds in parallel.  Returns a cudaError_t as int; -1 = shape not in this unit.
int jacobi_tile_launch_inst_a(int R, int NW, int mode, const TileArgs& A, int num_sms, bool pdl, cudaStream_t st);
int jacobi_tile_launch_inst_b(int R, int NW, int mode, const TileArgs& A, int num_sms, bool pdl, cudaStream_t st);
int jacobi_tile_launch_inst_c(int R, int NW, int mode, const TileArgs& A, int num_sms, bool pdl, cudaStream_t st);
int jacobi_tile_launch_inst_d(int R, int NW, int mode, const TileArgs& A, int num_sms, bool pdl, cudaStream_t st);

}  // namespace fruid
