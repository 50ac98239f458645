// kernels_vec.cuh -- float4-per-thread versions of the single-pass operators (add_source,
// divergence, gradient-subtract, advection).  One thread owns four consecutive cells of one
// row, so every global access is a 16-byte vector and a warp covers one 512-byte row
// segment.  Boundary cells are written by the thread that owns the adjacent interior cell
// (set_bnd folded into the producer, src/lib.rs:27-44).  Requires nx >= 1, ny >= 1,
// pitch % 4 == 0 and 16-byte aligned bases (true for every library-owned field).
#pragma once
#include "common.cuh"
#include "kernels_basic.cuh"
#include "slab_comm.cuh"

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
    pdl_enter();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4* X1 = reinterpret_cast<float4*>(x1);
    dbg_tail_delay(threadIdx.x >> 5);
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
    pdl_enter();
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
    dbg_tail_delay(threadIdx.y);
    store_row4_bnd(div, g, B_POSITIVE, gx0, gy, d);
    if (p) store_row4_bnd(p, g, B_POSITIVE, gx0, gy, make_float4(0.f, 0.f, 0.f, 0.f));
}

// Second half of project (src/lib.rs:164-168): u += gx*(p[i+1,j]-p[i-1,j]); v += gy*(p[i,j+1]-p[i,j-1]);
// set_bnd(NegX,u); set_bnd(NegY,v).  In place; each thread touches only its own cells.
__global__ void k_gradient4(float* __restrict__ u, float* __restrict__ v, const float* __restrict__ p, Geom g, float gxs,
                            float gys) {
    pdl_enter();
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
    dbg_tail_delay(threadIdx.y);
    store_row4_bnd(u, g, B_NEGX, gx0, gy, un);
    store_row4_bnd(v, g, B_NEGY, gx0, gy, vn);
}

// Where the rows of a field live when it is y-slab decomposed over several GPUs: rank r owns global
// rows [own0[r], own0[r+1]).  base[f][r] is rank r's copy of gather field f (a peer-mapped pointer,
// NVLink) SHIFTED so that it is indexed by GLOBAL row: row gj of rank r sits at base[f][r] + gj*pitch.
// A gather tap at global row gj reads the copy of its OWNER.
struct SlabTable {
    int world;                 // 1 = single GPU
    int me;
    int my_lo, my_hi;          // my owned global rows [my_lo, my_hi): taps inside use the local pointer
    int own0[9];
    const float* base[2][8];   // [field][rank], global-row indexed
    int* block_flag;           // [nblocks]: pass 1 marks blocks that still have remote cells ...
    int* block_list;           // ... and queues them here: [0] = count, [1..] = block ids (pass 2's work list)
    int grid_x;                // blocks per row of the pass-1 grid
    const int* src_done;       // my SlabFlags::src_done[]: src_done[r] >= event <=> rank r's fields are final
    int* error;
    int event;
    SlabFlags* flags[8];       // every rank's flag block (peer-mapped): pass 1 broadcasts src_done, pass 2 adv_done
    int* pass2_done;           // block counter of pass 2 (self-resetting)
};
__device__ __forceinline__ int slab_owner(const SlabTable& t, int gj) {
    int r = 0;  // branch-free: number of slab cuts at or below gj
#pragma unroll
    for (int k = 1; k < 8; ++k) r += (k < t.world && gj >= t.own0[k]) ? 1 : 0;
    return r;
}
// Row gj of gather field f as its owner holds it (local pointer for my own rows).
__device__ __forceinline__ const float* slab_row(const SlabTable& t, int f, const float* d0, int gj, size_t pitch) {
    const float* base = (gj >= t.my_lo && gj < t.my_hi) ? d0 : t.base[f][slab_owner(t, gj)];
    return base + (size_t)gj * pitch;
}
// Bilinear gather of one cell from local rows.  d0 is indexed by global row (on one GPU: the field).
__device__ __forceinline__ float gather_local(const float* __restrict__ d0, const Geom& g, const Trace& t) {
    const float* r0 = d0 + (size_t)t.j0 * g.pitch + t.i0;
    const float a00 = __ldg(r0), a10 = __ldg(r0 + 1), a01 = __ldg(r0 + g.pitch), a11 = __ldg(r0 + g.pitch + 1);
    return mixf(mixf(a00, a01, t.t1), mixf(a10, a11, t.t1), t.s1);
}
// Four cells whose tap rows may belong to other ranks: all sixteen addresses first, then all sixteen
// loads back to back (each is an NVLink round trip; serialising them would cost ~16x the latency).
__device__ __forceinline__ float4 gather_remote4(const SlabTable& tab, int f, const float* d0, const Geom& g, const Trace (&t)[4]) {
    const float* p[4][2];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        p[c][0] = slab_row(tab, f, d0, t[c].j0, g.pitch) + t[c].i0;
        p[c][1] = slab_row(tab, f, d0, t[c].j0 + 1, g.pitch) + t[c].i0;
    }
    float a[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {  // peer memory: volatile loads, no stale L1 lines
        a[c][0] = __ldcv(p[c][0]); a[c][1] = __ldcv(p[c][1]); a[c][2] = __ldcv(p[c][0] + 1); a[c][3] = __ldcv(p[c][1] + 1);
    }
    float r[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) r[c] = mixf(mixf(a[c][0], a[c][1], t[c].t1), mixf(a[c][2], a[c][3], t[c].t1), t[c].s1);
    return make_float4(r[0], r[1], r[2], r[3]);
}

// advect (src/lib.rs:84-126) for NF fields sharing one back-trace, four cells per thread; block
// (bx, by) covers 128 columns x 8 rows.  Output rows: local rows [ly0, ly1) (all interior rows on one
// GPU; the owned rows of a slab).  a.d0[f] must be indexed by GLOBAL row (slab mode: local pointer
// minus gy0 rows).
//   PASS 0: single GPU.
//   PASS 1: slab, common case -- threads whose eight tap rows are all mine do the work exactly like
//           PASS 0; the others only queue their block for pass 2.
//   PASS 2: slab, queued blocks only -- the remaining threads gather from the OWNERS' memory over
//           NVLink, after the owners have published their fields (src_done).
template <int NF, int PASS>
__device__ __forceinline__ void advect4_block(const AdvectArgs& a, const float* __restrict__ u, const float* __restrict__ v,
                                              const Geom& g, float dt0, float dt1, float xmax, float ymax, int ly0, int ly1,
                                              const SlabTable& tab, int bx, int by) {
    const int gx0 = 4 * (bx * blockDim.x + threadIdx.x);
    const int ly = by * blockDim.y + threadIdx.y + ly0;
    if (gx0 >= g.w || ly >= ly1) return;
    const int gy = ly + g.gy0;  // global row: the back-trace and the gather use global coordinates
    const size_t k = (size_t)ly * g.pitch + gx0;
    const float4 uc = ld4(u + k), vc = ld4(v + k);
    Trace t[4];
    t[0] = backtrace(gx0 + 0, gy, uc.x, vc.x, dt0, dt1, xmax, ymax);
    t[1] = backtrace(gx0 + 1, gy, uc.y, vc.y, dt0, dt1, xmax, ymax);
    t[2] = backtrace(gx0 + 2, gy, uc.z, vc.z, dt0, dt1, xmax, ymax);
    t[3] = backtrace(gx0 + 3, gy, uc.w, vc.w, dt0, dt1, xmax, ymax);
    if (PASS != 0) {
        const int jmin = min(min(t[0].j0, t[1].j0), min(t[2].j0, t[3].j0));
        const int jmax = max(max(t[0].j0, t[1].j0), max(t[2].j0, t[3].j0));
        const bool local = (jmin >= tab.my_lo) && (jmax + 1 < tab.my_hi);
        if (PASS == 1 && !local) {  // the first thread to mark the block also queues it for pass 2
            const int bid = by * tab.grid_x + bx;
            if (atomicExch(&tab.block_flag[bid], 1) == 0) tab.block_list[1 + atomicAdd(&tab.block_list[0], 1)] = bid;
            return;
        }
        if (PASS == 2 && local) return;  // done in pass 1
        if (PASS == 2) {  // taps beyond the neighbours' rows (large displacement): wait for those owners too
            const int r0 = slab_owner(tab, jmin), r1 = slab_owner(tab, jmax + 1);
            for (int r = r0; r <= r1; ++r)
                if (r < tab.me - 1 || r > tab.me + 1) spin_until(&tab.src_done[r], tab.event, tab.error);
        }
    }
    dbg_tail_delay(threadIdx.y);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        float4 r;
        if (PASS != 2) {
            r.x = gather_local(a.d0[f], g, t[0]);
            r.y = gather_local(a.d0[f], g, t[1]);
            r.z = gather_local(a.d0[f], g, t[2]);
            r.w = gather_local(a.d0[f], g, t[3]);
        } else {
            r = gather_remote4(tab, f, a.d0[f], g, t);
        }
        store_row4_bnd(a.d[f], g, a.b[f], gx0, ly, r);
    }
}

// PASS 0 / 1: one block per 128 x 8 cells.  PASS 2: a small persistent grid walks the list of blocks
// that pass 1 queued (typically the one or two block rows next to a slab cut).
template <int NF, int PASS>
__global__ void k_advect4(AdvectArgs a, const float* __restrict__ u, const float* __restrict__ v, Geom g, float dt0,
                          float dt1, float xmax, float ymax, int ly0, int ly1, const __grid_constant__ SlabTable tab) {
    pdl_enter();
    if (PASS != 2) {
        // Folded "src_done" broadcast: the kernels that wrote the gathered fields precede this launch in
        // stream order, so they are final; tell every rank (nobody waits here).
        if (PASS == 1 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.y == 0 && (int)threadIdx.x < tab.world)
            st_release_sys(&tab.flags[threadIdx.x]->src_done[tab.me], tab.event);
        advect4_block<NF, PASS>(a, u, v, g, dt0, dt1, xmax, ymax, ly0, ly1, tab, blockIdx.x, blockIdx.y);
        return;
    }
    const int count = tab.block_list[0];
    if (count == 0) {  // nothing was gathered remotely: "adv_done" can be broadcast right away
        if (blockIdx.x == 0 && threadIdx.y == 0 && (int)threadIdx.x < tab.world)
            st_release_sys(&tab.flags[threadIdx.x]->adv_done[tab.me], tab.event);
        return;
    }
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        // One thread waits (system-scope acquire) for the two NEIGHBOUR owners on behalf of the block --
        // a system fence per thread would serialise; farther owners are handled per thread.
        if (tab.me > 0) spin_until(&tab.src_done[tab.me - 1], tab.event, tab.error);
        if (tab.me + 1 < tab.world) spin_until(&tab.src_done[tab.me + 1], tab.event, tab.error);
    }
    __syncthreads();
    for (int i = blockIdx.x; i < count; i += gridDim.x) {
        const int bid = tab.block_list[1 + i];
        if (threadIdx.x == 0 && threadIdx.y == 0) tab.block_flag[bid] = 0;
        advect4_block<NF, 2>(a, u, v, g, dt0, dt1, xmax, ymax, ly0, ly1, tab, bid % tab.grid_x, bid / tab.grid_x);
    }
    // Folded "adv_done" broadcast by the last block to finish: this rank no longer reads anybody's fields.
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence();
        if (atomicAdd(tab.pass2_done, 1) == (int)gridDim.x - 1) {
            *tab.pass2_done = 0;
            tab.block_list[0] = 0;  // everybody has read the count: leave an empty work list for the next advect
            for (int r = 0; r < tab.world; ++r) st_release_sys(&tab.flags[r]->adv_done[tab.me], tab.event);
        }
    }
}

}  // namespace fruid
