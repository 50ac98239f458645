// slab_step.cuh -- the velocity and density step schedules of a y-slab decomposed field (one process per GPU).
#pragma once
#include "field_set.cuh"

using namespace fruid;

// ---- slab (multi-GPU) step schedules ---------------------------------------------------------
// Same operator sequence as op_vel_step / op_dens_step; every operator runs on all local rows it
// can (owned + halo), so a result stays valid a few rows beyond the owned ones and halo exchanges
// are needed only where the validity margin m (rows beyond the owned rows) runs out:
//   exchange -> m = 12;  T fused sweeps: m -> m - T;  divergence / gradient: m -> m - 1.
// Advection gathers from the OWNER of each tap through peer memory, between two all-rank
// barriers (everyone has finished writing / reading the gathered fields).
static int project_slab(FieldSet& S, float*& u, float*& v, float*& p, float*& div, int T) {
    const Geom& g = S.g;
    cudaStream_t st = S.stream;
    const float sx = -0.5f / (float)g.nx, sy = -0.5f / (float)g.ny;
    LAUNCH("k_divergence4", st, launch_pdl(true, k_divergence4, vec_grid(g), vec_block(), 0, st, div, (float*)nullptr, u, v, g, sx, sy));  // m: HALO -> HALO-1
    SlabSolve ss{SLAB_HALO, SLAB_HALO - 1, 1, [&S](std::initializer_list<float*> f) { return S.exchange(f); }};
    TRY(op_lin_solve(B_POSITIVE, p, div, S.buf[4], g, 1.0f, 4.0f, S.iters, T, st, true, false, 0.f, nullptr, &ss));
    const float gx = -0.5f * (float)g.nx, gy = -0.5f * (float)g.ny;
    LAUNCH("k_gradient4", st, launch_pdl(true, k_gradient4, vec_grid(g), vec_block(), 0, st, u, v, p, g, gx, gy));  // needs p one row out
    return 0;
}

static int op_vel_step_slab(FieldSet& S, float visc, float dt) {
    TRY(S.need_peers());
    const Geom& g = S.g;
    cudaStream_t st = S.stream;
    const int T = S.slab_fused();
    float *&u = S.buf[0], *&v = S.buf[1], *&u0 = S.buf[2], *&v0 = S.buf[3], *&s = S.buf[4];
    float* const s2 = S.buf[5];
    const auto xch = [&S](std::initializer_list<float*> f) { return S.exchange(f); };
    const float a = ((dt * visc) * (float)g.nx) * (float)g.ny, c = 1.0f + 4.0f * a;
    TRY(S.exchange({u, v, u0, v0}));                                                   // margins = HALO everywhere
    {   // add_source folded into the first block of each diffuse (lib.rs:190-197)
        SlabSolve su{SLAB_HALO, SLAB_HALO, 0, xch};
        TRY(op_lin_solve(B_NEGX, u0, u, s, g, a, c, S.iters, T, st, false, true, dt, s2, &su));
        SlabSolve sv{SLAB_HALO, SLAB_HALO, 0, xch};
        TRY(op_lin_solve(B_NEGY, v0, v, s, g, a, c, S.iters, T, st, false, true, dt, s2, &sv));
    }
    TRY(S.exchange({u0, v0}));
    TRY(S.wait_readers());                     // density steps of the previous scene step have read u, v
    TRY(project_slab(S, u0, v0, u, v, T));                                             // lib.rs:199 (overwrites u, v)
    S.slab.ev_adv += 1;                        // advect pass 1 broadcasts "my u0, v0 are final" (src_done)
    {
        SlabTable tab{};
        TRY(S.fill_table(tab, 0, u0));
        TRY(S.fill_table(tab, 1, v0));
        float* d[2] = {u, v};
        const float* d0[2] = {S.global_rows(u0), S.global_rows(v0)};
        const int b[2] = {B_NEGX, B_NEGY};
        const int ly0 = std::max(S.own_lo(), 1 - g.gy0), ly1 = std::min(S.own_hi(), g.gh - 1 - g.gy0);
        TRY(op_advect(2, d, d0, b, u0, v0, g, dt, st, &tab, ly0, ly1));                // lib.rs:204-205
    }
    // advect pass 2 broadcast "I have stopped reading everybody's u0, v0" (adv_done); the exchange below also
    // waits until everybody has, because the projection overwrites u0, v0 (as p, div)
    TRY(S.exchange({u, v}, S.slab.ev_adv));
    return project_slab(S, u, v, u0, v0, T);                                           // lib.rs:207
}

// `vel_ready` (recorded on the velocity handle's stream after its step) is awaited only right before
// the advect: the halo exchange and the diffusion solve do not read the velocity, so they (and the
// waiting inside the exchange) overlap the tail of the velocity step running on the other stream.
static int op_dens_step_slab(FieldSet& S, const float* u, const float* v, float diff, float dt, cudaEvent_t vel_ready) {
    TRY(S.need_peers());
    const Geom& g = S.g;
    cudaStream_t st = S.stream;
    const int T = S.slab_fused();
    float *&x = S.buf[0], *&x0 = S.buf[1], *&s = S.buf[2];
    float* const s2 = S.buf[3];
    const float a = ((dt * diff) * (float)g.nx) * (float)g.ny, c = 1.0f + 4.0f * a;
    TRY(S.exchange({x, x0}, S.slab.ev_adv));   // ... and everybody has finished the previous advect's reads of x0
    SlabSolve sd{SLAB_HALO, SLAB_HALO, 0, [&S](std::initializer_list<float*> f) { return S.exchange(f); }};
    TRY(op_lin_solve(B_POSITIVE, x0, x, s, g, a, c, S.iters, T, st, false, true, dt, s2, &sd));  // lib.rs:172,175
    S.slab.ev_adv += 1;                        // advect pass 1 broadcasts "my x0 is final"
    CU(stream_wait(st, vel_ready)); // u, v of this step
    SlabTable tab{};
    TRY(S.fill_table(tab, 0, x0));
    float* d[1] = {x};
    const float* d0[1] = {S.global_rows(x0)};
    const int b[1] = {B_POSITIVE};
    const int ly0 = std::max(S.own_lo(), 1 - g.gy0), ly1 = std::min(S.own_hi(), g.gh - 1 - g.gy0);
    return op_advect(1, d, d0, b, u, v, g, dt, st, &tab, ly0, ly1);                     // lib.rs:178 (pass 2 broadcasts adv_done)
}
