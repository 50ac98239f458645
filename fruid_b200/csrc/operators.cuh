// operators.cuh -- launchers of the solver's operators on device buffers, one per function of the reference
// (src/lib.rs:5-208): add_source, set_bnd, lin_solve, diffuse, advect, project, vel_step, dens_step.
#pragma once
#include <functional>
#include <initializer_list>
#include "host_runtime.cuh"
#include "kernels_basic.cuh"
#include "kernels_serial.cuh"
#include "kernels_vec.cuh"
#include "jacobi_tile.cuh"
#include "advect_tile.cuh"
#include "slab_comm.cuh"

using namespace fruid;

// --------------------------------------------------------------------------------------
// operator launchers on device buffers
// --------------------------------------------------------------------------------------
static inline size_t pitch_for(size_t w) { return (w + 31) / 32 * 32; }
static inline bool degenerate(const Geom& g) { return g.nx < 1 || g.ny < 1; }  // empty interior (w == 2 or h == 2)
static inline dim3 cell_block() { return dim3(128, 2); }
static inline dim3 cell_grid(int cx, int cy) { return dim3((cx + 127) / 128, (cy + 1) / 2); }
// float4-per-thread kernels: a warp covers 128 columns of one interior row
static inline dim3 vec_block() { return dim3(32, 8); }
static inline dim3 vec_grid(const Geom& g) { return dim3(((g.w + 3) / 4 + 31) / 32, (g.h - 2 + 7) / 8); }

static int op_add_source(float* x1, const float* s1, float* x2, const float* s2, const Geom& g, float dt,
                         cudaStream_t st) {
    const size_t n4 = g.pitch * (size_t)g.h / 4;
    LAUNCH("k_add_source4", st, launch_pdl(true, k_add_source4, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, st, x1, s1, x2, s2, n4, dt));
    return 0;
}

static int op_set_bnd(int b, float* x, const Geom& g, cudaStream_t st) {
    if (degenerate(g)) {
        LAUNCH("k_serial_set_bnd", st, k_serial_set_bnd<<<1, 1, 0, st>>>(x, g, b));
        return 0;
    }
    LAUNCH("k_set_bnd", st, k_set_bnd<<<cell_grid(g.nx, g.ny), cell_block(), 0, st>>>(x, g, b));
    return 0;
}

// lin_solve, src/lib.rs:46-71.  x / scratch ping-pong like the re-bound references at
// lib.rs:49-50,68.  `x` and `scratch` are passed by reference: on return `x` names the
// buffer holding the newest iterate (the two are exchanged when the number of launches
// is odd), so the caller's field roles follow the data and nothing is copied.
// fused = sweeps per launch (1 = k_jacobi1; >1 = register-tile kernel of jacobi_tile.cuh).
static inline int resolve_fused(const Geom& g, int iters, int fused) {
    if (fused <= 0) fused = jacobi_tile_default_fused(g, iters);
    return fused > iters ? iters : fused;
}
// Slab mode: validity margins (rows beyond the owned rows that hold correct values) of the
// iterate and of the right-hand side, the margin the caller needs on return, and the halo
// exchange to call when a block would run out of margin.  T fused sweeps cost the iterate T
// rows of margin; the result also needs the right-hand side T-1 rows further out.
struct SlabSolve {
    int m_x, m_x0, need_final;
    std::function<int(std::initializer_list<float*>)> exchange;
};
static int op_lin_solve(int b, float*& x, const float* x0, float*& scratch, const Geom& g, float a, float c,
                        int iters, int fused, cudaStream_t st, bool x_is_zero = false, bool fuse_src = false,
                        float src_dt = 0.f, float* z_tmp = nullptr, SlabSolve* slab = nullptr) {
    if (iters < 2 || (iters & 1)) return fail(FRUID_ERR_ODD_ITERATIONS, "iterations must be even and >= 2, got %d", iters);
    if (degenerate(g)) {
        LAUNCH("k_serial_lin_solve", st, k_serial_lin_solve<<<1, 1, 0, st>>>(x, x0, scratch, g, a, c, b, iters));
        return 0;
    }
    fused = resolve_fused(g, iters, fused);
    if (fused == 1) {
        for (int it = 0; it < iters; ++it) {
            LAUNCH("k_jacobi1", st, k_jacobi1<<<cell_grid(g.nx, g.ny), cell_block(), 0, st>>>(scratch, x, x0, g, a, c, b));
            float* t = x; x = scratch; scratch = t;
        }
        return 0;
    }
    const int nblocks = (iters + fused - 1) / fused;
    int done = 0;
#ifdef FRUID_DEBUG
    // scratch and the add_source temporary are write-before-read (lib.rs:65,68): fill them with NaNs so that a
    // launch that reads them before its producer has finished yields NaNs instead of plausible stale values
    CU(cudaMemsetAsync(scratch, 0xFF, g.pitch * (size_t)g.h * sizeof(float), st));
    if (fuse_src && nblocks > 1 && z_tmp && !slab) CU(cudaMemsetAsync(z_tmp, 0xFF, g.pitch * (size_t)g.h * sizeof(float), st));
    pdl_chain_break(st);
#endif
    for (int k = 0; k < nblocks; ++k) {
        const int T = (iters - done + (nblocks - k) - 1) / (nblocks - k);  // spread evenly
        const bool fs = fuse_src && k == 0;  // first block also performs add_source(x0, x, dt) -> z_tmp
        const bool xz = x_is_zero && k == 0;
        const float* rhs = (fuse_src && k > 0) ? z_tmp : x0;
        if (slab) {  // refresh halo rows from the neighbours only when this block would run out of margin
            const int req = (k + 1 == nblocks) ? slab->need_final : 0;
            const bool xs = !xz && slab->m_x - T < req, zs = slab->m_x0 - (T - 1) < req;
            if (xs && zs) { TRY(slab->exchange({x, const_cast<float*>(rhs)})); slab->m_x = slab->m_x0 = SLAB_HALO; }
            else if (xs) { TRY(slab->exchange({x})); slab->m_x = SLAB_HALO; }
            else if (zs) { TRY(slab->exchange({const_cast<float*>(rhs)})); slab->m_x0 = SLAB_HALO; }
        }
        prof_pre(st);
        if (jacobi_tile_launch(scratch, x, rhs, g, a, c, b, T, xz ? 1 : 0, st, fs ? 1 : 0, src_dt,
                               (fs && nblocks > 1) ? z_tmp : nullptr) != 0)
            return fail(FRUID_ERR_BAD_ARG, "fused sweeps T=%d not supported by the current tile config", T);
        TRY(launched("k_jacobi_tile", st));
        done += T;
        float* t = x; x = scratch; scratch = t;
        if (slab) {  // rows beyond the owned ones that the block just produced correctly
            const int stored = SLAB_HALO - (T + (T & 1));  // the tile kernel stores no closer than H to a slab cut
            slab->m_x = std::min(std::min((xz ? SLAB_HALO : slab->m_x) - T, slab->m_x0 - (T - 1)), stored);
            if (fs) slab->m_x0 = std::min(slab->m_x0, stored);  // z_tmp is written on the stored rows only
        }
    }
    return 0;
}

// diffuse, src/lib.rs:73-78: a = ((dt*diff)*nx)*ny in f32, c = 1 + 4a.
static int op_diffuse(int b, float*& x, const float* x0, float*& scratch, const Geom& g, float diff, float dt,
                      int iters, int fused, cudaStream_t st) {
    const float a = ((dt * diff) * (float)g.nx) * (float)g.ny;
    const float c = 1.0f + 4.0f * a;
    return op_lin_solve(b, x, x0, scratch, g, a, c, iters, fused, st);
}

// add_source(x0, x, dt) followed by diffuse(b, x, x0, ...) (lib.rs:172-175, 190-197: the source
// field is also the initial iterate of the solve).  With the fused solver the add is folded into
// the first block (registers) and x0 itself is left untouched -- its updated value would be dead:
// the callers overwrite x0 right after (project zeroes it as p, lib.rs:152; advect rewrites it).
static int op_add_source_diffuse(int b, float*& x, float* x0, float*& scratch, float* z_tmp, const Geom& g, float diff,
                                 float dt, int iters, int fused, cudaStream_t st) {
    const float a = ((dt * diff) * (float)g.nx) * (float)g.ny;
    const float c = 1.0f + 4.0f * a;
    if (degenerate(g) || resolve_fused(g, iters, fused) <= 1 || !z_tmp) {
        TRY(op_add_source(x0, x, nullptr, nullptr, g, dt, st));
        return op_lin_solve(b, x, x0, scratch, g, a, c, iters, fused, st);
    }
    return op_lin_solve(b, x, x0, scratch, g, a, c, iters, fused, st, false, true, dt, z_tmp);
}

// Which advect kernel runs on one GPU (fruid_set_advect_tile): 0 = k_advect4 (global gathers), 1 = k_advect_tile
// (gathered fields staged in shared memory), -1 = by size: the tiled kernel from ~1.5 M cells up -- it gives a thread
// 16 cells instead of 4, which only pays once there are several blocks per SM.
static std::atomic<int> g_advect_tile{-1};
static inline bool advect_use_tile(const Geom& g) {
    const int m = g_advect_tile.load(std::memory_order_relaxed);
    if (m >= 0) return m != 0;
    return (size_t)g.w * (size_t)g.h >= (size_t)1200 * 1200;
}

// advect, src/lib.rs:84-126, for 1..4 fields sharing the back-trace (slab mode: 1 or 2).  `tab` (slab mode) says where
// the rows of the gathered fields live on the other ranks; d0[] must then be indexed by GLOBAL row
// and [ly0, ly1) are the local rows to produce (default: every interior row).
static int op_advect(int nf, float* const* d, const float* const* d0, const int* b, const float* u, const float* v,
                     const Geom& g, float dt, cudaStream_t st, const SlabTable* tab = nullptr, int ly0 = -1, int ly1 = -1) {
    const float dt0 = dt * (float)g.nx, dt1 = dt * (float)g.ny;
    const float xmax = (float)g.nx + 0.5f, ymax = (float)g.ny + 0.5f;
    if (degenerate(g)) {
        for (int f = 0; f < nf; ++f) {
            // interior loops are empty: only the trailing set_bnd (lib.rs:125) acts
            LAUNCH("k_serial_set_bnd", st, k_serial_set_bnd<<<1, 1, 0, st>>>(d[f], g, b[f]));
        }
        return 0;
    }
    AdvectArgs a{};
    for (int f = 0; f < nf; ++f) { a.d[f] = d[f]; a.d0[f] = d0[f]; a.b[f] = b[f]; }
    SlabTable single{};
    single.world = 1;
    const bool multi = tab && tab->world > 1;
    if (!multi) tab = &single;
    if (ly0 < 0) { ly0 = 1; ly1 = g.h - 1; }
    if (ly1 <= ly0) return 0;
    const dim3 grid1(((g.w + 3) / 4 + 31) / 32, (ly1 - ly0 + 7) / 8);
    SlabTable T = *tab;
    T.grid_x = (int)grid1.x;
#define FRUID_ADVECT(NF, PASS, name) \
    LAUNCH(name, st, launch_pdl(PASS != 2, k_advect4<NF, PASS>, (PASS == 2 ? dim3(296) : grid1), vec_block(), 0, st, a, u, v, g, dt0, dt1, xmax, ymax, ly0, ly1, T))
    if (nf < 1 || nf > ADVECT_MAX_FIELDS || (multi && nf > 2)) return fail(FRUID_ERR_BAD_ARG, "advect: unsupported field count %d", nf);
    // TMA needs 16-byte aligned bases and row pitches (true for library-owned fields; the fruid_dev_* entry points take
    // the caller's arrays as they come -- those fall back to the gather kernel instead of failing)
    bool tma_ok = (g.pitch % 4 == 0) && ((uintptr_t)u % 16 == 0) && ((uintptr_t)v % 16 == 0);
    for (int f = 0; f < nf; ++f) tma_ok = tma_ok && ((uintptr_t)d0[f] % 16 == 0);
    if (tma_ok && advect_use_tile(g) && (multi ? nf <= 2 : (ly0 == 1 && ly1 == g.h - 1))) {
        const bool selfadv = (nf == 2) && (d0[0] + (size_t)g.gy0 * g.pitch == u) && (d0[1] + (size_t)g.gy0 * g.pitch == v);  // lib.rs:204-205
        int r;
        prof_pre(st);
#define FRUID_ADVT(NF, UV, PASS) advect_tile_launch<NF, UV, PASS>(a, u, v, g, dt0, dt1, xmax, ymax, ly0, ly1, T, st)
        if (!multi) {
            switch (nf) {
                case 1: r = FRUID_ADVT(1, ADV_UV_TMA, 0); break;
                case 2: r = selfadv ? FRUID_ADVT(2, ADV_UV_SELF, 0) : FRUID_ADVT(2, ADV_UV_TMA, 0); break;
                case 3: r = FRUID_ADVT(3, ADV_UV_LDG, 0); break;
                default: r = FRUID_ADVT(4, ADV_UV_LDG, 0); break;
            }
        } else {
            if (nf == 1) r = FRUID_ADVT(1, ADV_UV_TMA, 1);
            else r = selfadv ? FRUID_ADVT(2, ADV_UV_SELF, 1) : FRUID_ADVT(2, ADV_UV_TMA, 1);
        }
#undef FRUID_ADVT
        if (r == -4) return fail(FRUID_ERR_CUDA, "advect: no tensor map for the gathered fields");
        static const char* const names[] = {"", "k_advect_tile_x1", "k_advect_tile_x2", "k_advect_tile_x3", "k_advect_tile_x4"};
        TRY(launched(names[nf], st));
        if (multi) {  // pass 2: the quads that pass 1 queued, from the owners' memory
            if (nf == 2) { FRUID_ADVECT(2, 2, "k_advect4x2_remote"); }
            else { FRUID_ADVECT(1, 2, "k_advect4x1_remote"); }
        }
    } else if (!multi) {
        switch (nf) {
            case 1: FRUID_ADVECT(1, 0, "k_advect4x1"); break;
            case 2: FRUID_ADVECT(2, 0, "k_advect4x2"); break;
            case 3: FRUID_ADVECT(3, 0, "k_advect4x3"); break;
            default: FRUID_ADVECT(4, 0, "k_advect4x4"); break;
        }
    } else {  // pass 1: cells whose taps are all local; pass 2: the rest, from the owners' memory
        if (nf == 2) { FRUID_ADVECT(2, 1, "k_advect4x2"); FRUID_ADVECT(2, 2, "k_advect4x2_remote"); }
        else { FRUID_ADVECT(1, 1, "k_advect4x1"); FRUID_ADVECT(1, 2, "k_advect4x1_remote"); }
    }
#undef FRUID_ADVECT
    return 0;
}

// project, src/lib.rs:146-169.
static int op_project(float* u, float* v, float*& p, float* div, float*& scratch, const Geom& g, int iters, int fused,
                      cudaStream_t st) {
    if (degenerate(g)) {
        LAUNCH("k_serial_project", st, k_serial_project<<<1, 1, 0, st>>>(u, v, p, div, scratch, g, iters));
        return 0;
    }
    const float sx = -0.5f / (float)g.nx, sy = -0.5f / (float)g.ny;  // lib.rs:156-157
    // p is zeroed at lib.rs:152,160 (interior, then border/corners from the zero interior): with
    // the fused solver the zero field is not materialised, the first block is told instead.
    const bool p_implicit_zero = resolve_fused(g, iters, fused) > 1;
    LAUNCH("k_divergence4", st,
           launch_pdl(true, k_divergence4, vec_grid(g), vec_block(), 0, st, div, p_implicit_zero ? nullptr : p, u, v, g, sx, sy));
    TRY(op_lin_solve(B_POSITIVE, p, div, scratch, g, 1.0f, 4.0f, iters, fused, st, p_implicit_zero));  // lib.rs:162
    const float gx = -0.5f * (float)g.nx, gy = -0.5f * (float)g.ny;  // lib.rs:164-165
    LAUNCH("k_gradient4", st, launch_pdl(true, k_gradient4, vec_grid(g), vec_block(), 0, st, u, v, p, g, gx, gy));
    return 0;
}

// vel_step, src/lib.rs:181-208 (argument roles alternate exactly as in the reference;
// no SWAPs -- lib.rs:193,196,201-202 are commented out there).
// The two diffuses are independent (lib.rs:194 touches u*, lib.rs:197 v*): given a second stream and a second pair of
// scratch buffers the v solve runs beside the u solve (on small grids a lone solve leaves most SMs idle).
struct VelFork {
    float*& scratch_v;
    float* tmp_v;
    cudaStream_t st2;
    cudaEvent_t fork, join;
};
static int op_vel_step(float*& u, float*& v, float*& u0, float*& v0, float*& s, float* s2, const Geom& g, float visc,
                       float dt, int iters, int fused, cudaStream_t st, VelFork* fk = nullptr) {
    if (fk) {
        CU(cudaEventRecord(fk->fork, st));
        CU(stream_wait(fk->st2, fk->fork));
        TRY(op_add_source_diffuse(B_NEGY, v0, v, fk->scratch_v, fk->tmp_v, g, visc, dt, iters, fused, fk->st2));  // lib.rs:191,197
        CU(cudaEventRecord(fk->join, fk->st2));
        TRY(op_add_source_diffuse(B_NEGX, u0, u, s, s2, g, visc, dt, iters, fused, st));                          // lib.rs:190,194
        CU(stream_wait(st, fk->join));
    } else {
        TRY(op_add_source_diffuse(B_NEGX, u0, u, s, s2, g, visc, dt, iters, fused, st));  // lib.rs:190,194
        TRY(op_add_source_diffuse(B_NEGY, v0, v, s, s2, g, visc, dt, iters, fused, st));  // lib.rs:191,197
    }
    TRY(op_project(u0, v0, u, v, s, g, iters, fused, st));                  // lib.rs:199
    float* d[2] = {u, v};
    const float* d0[2] = {u0, v0};
    const int b[2] = {B_NEGX, B_NEGY};
    TRY(op_advect(2, d, d0, b, u0, v0, g, dt, st));                         // lib.rs:204-205
    return op_project(u, v, u0, v0, s, g, iters, fused, st);                // lib.rs:207
}

// dens_step, src/lib.rs:171-179.
static int op_dens_step(float*& x, float*& x0, const float* u, const float* v, float*& s, float* s2, const Geom& g,
                        float diff, float dt, int iters, int fused, cudaStream_t st) {
    TRY(op_add_source_diffuse(B_POSITIVE, x0, x, s, s2, g, diff, dt, iters, fused, st));  // lib.rs:172,175
    float* d[1] = {x};
    const float* d0[1] = {x0};
    const int b[1] = {B_POSITIVE};
    return op_advect(1, d, d0, b, u, v, g, dt, st);                         // lib.rs:178
}
