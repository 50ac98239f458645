// fruid_abi.cu -- the C ABI of include/fruid_b200.h: device twins of FluidSim /
// DensitySim (reference src/lib.rs:210-294), the step schedules vel_step / dens_step
// (src/lib.rs:171-208) and the per-operator entry points.  No CPU fallback anywhere:
// every entry point either runs the CUDA kernels or returns an error.
#include "../../include/fruid_b200.h"

#include <algorithm>
#include <atomic>
#include <functional>
#include <initializer_list>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>
#include "field_set.cuh"
#include "slab_step.cuh"
#include "draw.cuh"

using namespace fruid;

extern "C" int fruid_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
    return 0;
}
extern "C" int fruid_profile_report(fruid_kernel_stat* out, size_t cap, size_t* n_out) {
    if (!n_out) return fail(FRUID_ERR_BAD_ARG, "null n_out");
    CU(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    size_t n = 0;
    for (const ProfRec& r : g_prof) {
        if (r.name && r.e1) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
                size_t k = 0;
                for (; k < n; ++k) if (strncmp(out[k].name, r.name, sizeof out[k].name) == 0) break;
                if (k == n && n < cap) { memset(&out[n], 0, sizeof out[n]); strncpy(out[n].name, r.name, sizeof out[n].name - 1); ++n; }
                if (k < cap) { out[k].launches += 1; out[k].total_ms += ms; }
            }
        }
        if (r.e0) g_prof_pool.push_back(r.e0);
        if (r.e1) g_prof_pool.push_back(r.e1);
    }
    g_prof.clear();
    *n_out = n;
    return 0;
}

extern "C" const char* fruid_last_error(void) { return g_err; }
extern "C" uint64_t fruid_kernel_launch_count(void) { return g_launches.load(); }
extern "C" const char* fruid_build_info(void) {
#ifdef FRUID_DEBUG
    return "fruid_b200 0.2 (sm_100a, -fmad=false, IEEE f32 rn) FRUID_DEBUG";
#else
    return "fruid_b200 0.2 (sm_100a, -fmad=false, IEEE f32 rn)";
#endif
}

static bool g_vel_fork = true;      // u and v diffuse on two streams
// Density add_source + diffuse beside the velocity step still in flight (only the advect waits for u, v): on by
// default (250^2 scene step 85 -> 70 us, 4096^2 1.15 -> 1.135 ms).  FRUID_B200_DENS_OVERLAP=0 / fruid_debug_set_schedule
// select one captured graph of the whole density step after the velocity step instead.  (Round 1 kept this off
// because of an "intermittent mismatch" that turned out to be a use-after-free in the TEST; every
// schedule is now checked on the FRUID_DEBUG build with delayed stores, tests/test_ordering_gpu.py.)
static bool g_dens_overlap = true;

// ---- programmatic launch chains (declared in common.cuh) ------------------------------------------------
namespace fruid {
PdlState g_pdl;
int g_dbg_tail_delay_host = 0;
namespace {
struct ChainSlot { cudaStream_t st; bool open; };
thread_local ChainSlot t_chain[16];
thread_local int t_chain_n = 0;
}  // namespace
bool pdl_chain_continue(cudaStream_t st) {
    for (int i = 0; i < t_chain_n; ++i)
        if (t_chain[i].st == st) { const bool was = t_chain[i].open; t_chain[i].open = true; return was; }
    if (t_chain_n < 16) t_chain[t_chain_n++] = ChainSlot{st, true};
    return false;  // (more than 16 streams in one call: the extra ones simply never chain)
}
void pdl_chain_break(cudaStream_t st) {
    for (int i = 0; i < t_chain_n; ++i)
        if (t_chain[i].st == st) { t_chain[i].open = false; return; }
}
void pdl_chain_reset() { t_chain_n = 0; }
int pdl_dev_flag_sync(int* dev_state, const void* symbol) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    const int v = g_pdl.enabled ? 1 : 0;
    // every device this process has a context on reads its own copy of the flag
    int ndev = 0, cur = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || cudaGetDevice(&cur) != cudaSuccess) return -1;
    (void)ndev;
    if (cudaMemcpyToSymbol(symbol, &v, sizeof v) != cudaSuccess) { cudaGetLastError(); return -1; }
    *dev_state = v;
    return 0;
}
}  // namespace fruid

namespace { struct PdlEnv { PdlEnv() {
    if (const char* e = getenv("FRUID_B200_LAUNCH_OVERLAP")) g_pdl.enabled = atoi(e) != 0;
    if (const char* e = getenv("FRUID_B200_PDL_UNCHAINED")) g_pdl.unsafe = atoi(e) != 0;
    if (const char* e = getenv("FRUID_B200_VEL_FORK")) g_vel_fork = atoi(e) != 0;
    if (const char* e = getenv("FRUID_B200_DENS_OVERLAP")) g_dens_overlap = atoi(e) != 0;
    if (const char* e = getenv("FRUID_B200_ADVECT_TILE")) g_advect_tile.store(atoi(e));
} } g_pdl_env; }
// Tuning / debugging knob: programmatic dependent launch between the kernels of a step (default on).
extern "C" int fruid_set_launch_overlap(int on) { g_pdl.enabled = on != 0; g_cfg_epoch.fetch_add(1); return 0; }
// Tuning knob: advect kernel on one GPU (0 = global gathers, 1 = shared-memory window, -1 = by grid size).
extern "C" int fruid_set_advect_tile(int mode) { g_advect_tile.store(mode < 0 ? -1 : (mode != 0)); g_cfg_epoch.fetch_add(1); return 0; }
// Tuning / debugging knob: replay steps as CUDA graphs (default on).
extern "C" int fruid_set_graphs_enabled(int on) { g_graphs_enabled = on != 0; g_cfg_epoch.fetch_add(1); return 0; }
// Debug knobs of the ordering tests.  bit 0: density diffuse beside the velocity step (default on), bit 1: u / v
// diffuse fork (default on), bit 2: round-1 launch attribution (programmatic attribute on EVERY launch; unsafe).
extern "C" int fruid_debug_set_schedule(int mask) {
    g_dens_overlap = (mask & 1) != 0; g_vel_fork = (mask & 2) != 0; g_pdl.unsafe = (mask & 4) != 0;
    g_cfg_epoch.fetch_add(1);
    return 0;
}
// Diagnostics: how the library divides by `c` (the cached outcome of the exhaustive reciprocal check): out = {seen, mode}.
extern "C" int fruid_debug_recip_mode(float c, int* out, float* zhzl) {
    if (!out || !zhzl) return fail(FRUID_ERR_BAD_ARG, "null out pointer");
    uint32_t key; memcpy(&key, &c, 4);
    std::lock_guard<std::mutex> lk(g_recip_mu);
    auto it = g_recip.find(key);
    out[0] = it != g_recip.end(); out[1] = out[0] ? it->second.mode : -1;
    zhzl[0] = out[0] ? it->second.zh : 0.f; zhzl[1] = out[0] ? it->second.zl : 0.f;
    return 0;
}
// FRUID_DEBUG builds: the lowest warps of every block wait `ns` before storing (0 = off).  Release builds: no-op, returns 1.
extern "C" int fruid_debug_set_tail_delay(int ns) {
#ifdef FRUID_DEBUG
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyToSymbol(g_dbg_tail_delay_ns, &ns, sizeof ns));
    g_dbg_tail_delay_host = ns;
    g_cfg_epoch.fetch_add(1);  // captured graphs bake the tile kernels' argument in
    return 0;
#else
    (void)ns;
    return 1;
#endif
}

// Live single-process registry of velocity handles: fruid_density_step_host looks its (u, v) pointers up here.
static std::mutex g_fluids_mu;
static std::vector<fruid_fluid*> g_fluids;
static std::atomic<uint64_t> g_velocity_uploads{0};

#define NEED(h) do { if (!(h)) return fail(FRUID_ERR_BAD_ARG, "null handle"); pdl_chain_reset(); CU(cudaSetDevice((h)->s.device)); } while (0)

extern "C" int fruid_fluid_create(size_t w, size_t h, fruid_fluid_t** out) {
    if (!out) return fail(FRUID_ERR_BAD_ARG, "null out pointer");
    *out = nullptr;
    fruid_fluid* f = new (std::nothrow) fruid_fluid();
    if (!f) return fail(FRUID_ERR_OOM, "host allocation failed");
    int r = f->s.init(w, h, 6, 0, 1, 2);  // + scratch and add_source temporary of the concurrent v diffuse
    if (r) { f->s.release(); delete f; return r; }
    { std::lock_guard<std::mutex> lk(g_fluids_mu); g_fluids.push_back(f); }
    *out = f;
    return 0;
}
// Multi-GPU: rank `rank` of `world` owns a y-slab of the global w x gh field (one process per GPU).
extern "C" int fruid_fluid_create_slab(size_t w, size_t gh, int rank, int world, fruid_fluid_t** out) {
    if (!out) return fail(FRUID_ERR_BAD_ARG, "null out pointer");
    *out = nullptr;
    fruid_fluid* f = new (std::nothrow) fruid_fluid();
    if (!f) return fail(FRUID_ERR_OOM, "host allocation failed");
    int r = f->s.init(w, gh, 6, rank, world);
    if (r) { f->s.release(); delete f; return r; }
    *out = f;
    return 0;
}
extern "C" int fruid_fluid_ipc_export(fruid_fluid_t* f, void* handle64) { NEED(f); return f->s.ipc_export(handle64); }
extern "C" int fruid_fluid_ipc_attach(fruid_fluid_t* f, int peer, const void* handle64) { NEED(f); return f->s.ipc_attach(peer, handle64); }
extern "C" int fruid_fluid_slab_rows(const fruid_fluid_t* f, size_t* y0, size_t* y1) {
    if (!f || !y0 || !y1) return fail(FRUID_ERR_BAD_ARG, "null pointer");
    *y0 = (size_t)f->s.slab.own0[f->s.slab.rank];
    *y1 = (size_t)f->s.slab.own0[f->s.slab.rank + 1];
    return 0;
}
// Diagnostics: n back-to-back halo exchanges of `nfields` fields / n all-rank barriers.
extern "C" int fruid_fluid_debug_exchange(fruid_fluid_t* f, int nfields, int n) {
    NEED(f);
    FieldSet& s = f->s;
    TRY(s.need_peers());
    for (int i = 0; i < n; ++i) {
        if (nfields >= 4) TRY(s.exchange({s.buf[0], s.buf[1], s.buf[2], s.buf[3]}));
        else if (nfields >= 2) TRY(s.exchange({s.buf[0], s.buf[1]}));
        else TRY(s.exchange({s.buf[0]}));
    }
    return 0;
}
extern "C" int fruid_fluid_debug_barrier(fruid_fluid_t* f, int n) {
    NEED(f);
    TRY(f->s.need_peers());
    for (int i = 0; i < n; ++i) TRY(f->s.barrier());
    return 0;
}
extern "C" int fruid_fluid_destroy(fruid_fluid_t* f) {
    if (!f) return 0;
    {
        std::lock_guard<std::mutex> lk(g_fluids_mu);
        g_fluids.erase(std::remove(g_fluids.begin(), g_fluids.end(), f), g_fluids.end());
    }
    cudaSetDevice(f->s.device);
    f->s.release();
    delete f;
    return 0;
}
extern "C" size_t fruid_fluid_width(const fruid_fluid_t* f) { return f ? (size_t)f->s.g.w : 0; }
extern "C" size_t fruid_fluid_height(const fruid_fluid_t* f) { return f ? (size_t)f->s.g.gh : 0; }
extern "C" int fruid_fluid_upload(fruid_fluid_t* f, int field, const float* host) { NEED(f); return f->s.upload(field, host); }
extern "C" int fruid_fluid_download(fruid_fluid_t* f, int field, float* host) {
    NEED(f);
    return f->s.download(field, host);
}
extern "C" int fruid_fluid_upload_async(fruid_fluid_t* f, int field, const float* host) { NEED(f); return f->s.upload_async(field, host); }
extern "C" int fruid_fluid_download_async(fruid_fluid_t* f, int field, float* host) { NEED(f); return f->s.download_async(field, host); }
extern "C" int fruid_fluid_inject(fruid_fluid_t* f, int field, size_t x, size_t y, float val) { NEED(f); return f->s.inject(field, x, y, val); }
extern "C" int fruid_fluid_set_iterations(fruid_fluid_t* f, int k) { NEED(f); return f->s.set_iters(k); }
extern "C" int fruid_fluid_set_fused_sweeps(fruid_fluid_t* f, int t) { NEED(f); return f->s.set_fused(t); }
// The caller's host arrays that shadow field `field` (the Array2D inside the crate-API mirror).
extern "C" int fruid_fluid_bind_mirror(fruid_fluid_t* f, int field, const float* host) {
    NEED(f);
    if (field < 0 || field >= f->s.n_user) return fail(FRUID_ERR_BAD_ARG, "bad field %d", field);
    std::lock_guard<std::mutex> lk(g_fluids_mu);
    f->s.mirror[field] = host;
    f->s.mirror_epoch[field] = 0;
    return 0;
}
extern "C" uint64_t fruid_host_velocity_uploads(void) { return g_velocity_uploads.load(); }
extern "C" int fruid_fluid_step(fruid_fluid_t* f, float dt, float visc) {
    NEED(f);
    FieldSet& s = f->s;
    s.touch();
    TRY(s.consume_uploads());
    if (s.is_slab()) return op_vel_step_slab(s, visc, dt);
    return run_step_graphed(s, dt, visc, nullptr, nullptr, [&]() {
        VelFork fk{s.buf[6], s.buf[7], s.stream2, s.ev_fork, s.ev_join};
        return op_vel_step(s.buf[0], s.buf[1], s.buf[2], s.buf[3], s.buf[4], s.buf[5], s.g, visc, dt, s.iters, s.fused, s.stream,
                           (s.stream2 && g_vel_fork) ? &fk : nullptr);
    });
}
extern "C" int fruid_fluid_step_n(fruid_fluid_t* f, float dt, float visc, int n) {
    NEED(f);
    if (n < 0) return fail(FRUID_ERR_BAD_ARG, "negative step count");
    for (int i = 0; i < n; ++i) TRY(fruid_fluid_step(f, dt, visc));
    return 0;
}
extern "C" void* fruid_fluid_stream(fruid_fluid_t* f) { return f ? (void*)f->s.stream : nullptr; }
extern "C" int fruid_fluid_sync(fruid_fluid_t* f) { NEED(f); CU(cudaStreamSynchronize(f->s.stream)); TRY(f->s.sync_copies()); CU(cudaGetLastError()); return f->s.check_peer_error(); }

extern "C" int fruid_density_create(size_t w, size_t h, fruid_density_t** out) {
    if (!out) return fail(FRUID_ERR_BAD_ARG, "null out pointer");
    *out = nullptr;
    fruid_density* d = new (std::nothrow) fruid_density();
    if (!d) return fail(FRUID_ERR_OOM, "host allocation failed");
    int r = d->s.init(w, h, 4);
    if (r) { d->s.release(); delete d; return r; }
    *out = d;
    return 0;
}
extern "C" int fruid_density_create_slab(size_t w, size_t gh, int rank, int world, fruid_density_t** out) {
    if (!out) return fail(FRUID_ERR_BAD_ARG, "null out pointer");
    *out = nullptr;
    fruid_density* d = new (std::nothrow) fruid_density();
    if (!d) return fail(FRUID_ERR_OOM, "host allocation failed");
    int r = d->s.init(w, gh, 4, rank, world);
    if (r) { d->s.release(); delete d; return r; }
    *out = d;
    return 0;
}
extern "C" int fruid_density_ipc_export(fruid_density_t* d, void* handle64) { NEED(d); return d->s.ipc_export(handle64); }
extern "C" int fruid_density_ipc_attach(fruid_density_t* d, int peer, const void* handle64) { NEED(d); return d->s.ipc_attach(peer, handle64); }
extern "C" int fruid_density_destroy(fruid_density_t* d) {
    if (!d) return 0;
    cudaSetDevice(d->s.device);
    d->s.release();
    delete d;
    return 0;
}
extern "C" int fruid_density_upload(fruid_density_t* d, int field, const float* host) { NEED(d); return d->s.upload(field, host); }
extern "C" int fruid_density_download(fruid_density_t* d, int field, float* host) {
    NEED(d);
    return d->s.download(field, host);
}
extern "C" int fruid_density_upload_async(fruid_density_t* d, int field, const float* host) { NEED(d); return d->s.upload_async(field, host); }
extern "C" int fruid_density_download_async(fruid_density_t* d, int field, float* host) { NEED(d); return d->s.download_async(field, host); }
extern "C" int fruid_density_inject(fruid_density_t* d, int field, size_t x, size_t y, float val) { NEED(d); return d->s.inject(field, x, y, val); }
// `field.data_mut().fill(val)` (src/main.rs:105-108) on the device.
static int fill_field(FieldSet& s, int field, float val) {
    if (field < 0 || field >= s.n_user) return fail(FRUID_ERR_BAD_ARG, "bad field %d", field);
    TRY(s.wait_readers());
    s.touch();
    if (val == 0.0f && !std::signbit(val)) {
        CU(cudaMemsetAsync(s.buf[field], 0, s.g.pitch * (size_t)s.g.h * sizeof(float), s.stream));
        return 0;
    }
    LAUNCH("k_fill", s.stream, k_fill<<<cell_grid(s.g.w, s.g.h), cell_block(), 0, s.stream>>>(s.buf[field], s.g, val));
    return 0;
}
extern "C" int fruid_density_fill(fruid_density_t* d, int field, float val) { NEED(d); return fill_field(d->s, field, val); }
extern "C" int fruid_fluid_fill(fruid_fluid_t* f, int field, float val) { NEED(f); return fill_field(f->s, field, val); }
extern "C" int fruid_density_set_iterations(fruid_density_t* d, int k) { NEED(d); return d->s.set_iters(k); }
extern "C" int fruid_density_set_fused_sweeps(fruid_density_t* d, int t) { NEED(d); return d->s.set_fused(t); }
extern "C" int fruid_density_step_dev(fruid_density_t* d, fruid_fluid_t* vel, float dt, float diff) {
    NEED(d);
    if (!vel) return fail(FRUID_ERR_BAD_ARG, "null velocity handle");
    FieldSet& s = d->s;
    FieldSet& vs = vel->s;
    if (vs.g.w != s.g.w || vs.g.gh != s.g.gh)
        return fail(FRUID_ERR_SIZE_MISMATCH, "density %dx%d vs velocity %dx%d", s.g.w, s.g.gh, vs.g.w, vs.g.gh);
    if (vs.device != s.device) return fail(FRUID_ERR_BAD_ARG, "density and velocity live on different devices");
    TRY(s.consume_uploads());
    if (s.is_slab() != vs.is_slab() || s.slab.world != vs.slab.world || s.slab.rank != vs.slab.rank)
        return fail(FRUID_ERR_SIZE_MISMATCH, "density and velocity handles are decomposed differently");
    // u, v are produced on vel's stream and consumed on ours; order both ways.  Only the advect needs them:
    // add_source + diffuse (lib.rs:172-175) start right away and run beside the velocity step still in flight.
    CU(cudaEventRecord(vs.ev, vs.stream));
    if (s.is_slab()) {
        TRY(op_dens_step_slab(s, vs.buf[0], vs.buf[1], diff, dt, vs.ev));
    } else if (!g_dens_overlap) {  // the whole step is one captured graph, after the velocity step
        CU(stream_wait(s.stream, vs.ev));
        TRY(run_step_graphed(s, dt, diff, vs.buf[0], vs.buf[1], [&]() {
            return op_dens_step(s.buf[0], s.buf[1], vs.buf[0], vs.buf[1], s.buf[2], s.buf[3], s.g, diff, dt, s.iters, s.fused, s.stream);
        }, vs.uid));
    } else {
        TRY(run_step_graphed(s, dt, diff, nullptr, nullptr, [&]() {
            return op_add_source_diffuse(B_POSITIVE, s.buf[1], s.buf[0], s.buf[2], s.buf[3], s.g, diff, dt, s.iters, s.fused, s.stream);
        }));
        CU(stream_wait(s.stream, vs.ev));
        float* dd[1] = {s.buf[0]};
        const float* d0[1] = {s.buf[1]};
        const int b[1] = {B_POSITIVE};
        TRY(op_advect(1, dd, d0, b, vs.buf[0], vs.buf[1], s.g, dt, s.stream));  // lib.rs:178
    }
    CU(cudaEventRecord(s.ev, s.stream));
    if (s.is_slab()) {
        if (std::find(vs.readers.begin(), vs.readers.end(), s.ev) == vs.readers.end()) vs.readers.push_back(s.ev);
    } else {
        CU(stream_wait(vs.stream, s.ev));
    }
    return 0;
}
// Several dye fields advanced by one velocity field (the four DensitySims of main.rs:114-118): every
// dye's add_source + diffuse runs on its own stream (concurrently on small grids; one captured graph with a
// fork and a join), then ONE advect launch moves up to four dyes with a shared back-trace (u, v are read once
// instead of once per dye).  Results are those of n separate fruid_density_step_dev calls, bit for bit.
static int dens_diffuse_batch_body(const std::vector<FieldSet*>& D, float dt, float diff) {
    FieldSet& L = *D[0];
    CU(cudaEventRecord(L.ev_fork, L.stream));
    for (size_t i = 0; i < D.size(); ++i) {
        FieldSet& s = *D[i];
        if (i) CU(stream_wait(s.stream, L.ev_fork));
        TRY(op_add_source_diffuse(B_POSITIVE, s.buf[1], s.buf[0], s.buf[2], s.buf[3], s.g, diff, dt, s.iters, s.fused, s.stream));  // lib.rs:172,175
        if (i) {
            CU(cudaEventRecord(s.ev_join, s.stream));
            CU(stream_wait(L.stream, s.ev_join));
        }
    }
    return 0;
}
static int dens_advect_batch(const std::vector<FieldSet*>& D, FieldSet& V, float dt) {
    FieldSet& L = *D[0];
    for (size_t i0 = 0; i0 < D.size(); i0 += ADVECT_MAX_FIELDS) {
        const int nf = (int)std::min<size_t>(ADVECT_MAX_FIELDS, D.size() - i0);
        float* d[ADVECT_MAX_FIELDS];
        const float* d0[ADVECT_MAX_FIELDS];
        int b[ADVECT_MAX_FIELDS];
        for (int f = 0; f < nf; ++f) { d[f] = D[i0 + f]->buf[0]; d0[f] = D[i0 + f]->buf[1]; b[f] = B_POSITIVE; }
        TRY(op_advect(nf, d, d0, b, V.buf[0], V.buf[1], L.g, dt, L.stream));  // lib.rs:178
    }
    return 0;
}
extern "C" int fruid_density_step_dev_batch(fruid_density_t* const* dyes, int n, fruid_fluid_t* vel, float dt, float diff) {
    if (!dyes || n < 0) return fail(FRUID_ERR_BAD_ARG, "null dye array or negative count");
    if (!vel) return fail(FRUID_ERR_BAD_ARG, "null velocity handle");
    if (n == 0) return 0;
    for (int i = 0; i < n; ++i) {
        if (!dyes[i]) return fail(FRUID_ERR_BAD_ARG, "null handle (dye %d)", i);
        for (int k = 0; k < i; ++k)
            if (dyes[k] == dyes[i]) return fail(FRUID_ERR_BAD_ARG, "dye %d given twice", i);
    }
    pdl_chain_reset();
    FieldSet& L = dyes[0]->s;
    FieldSet& V = vel->s;
    bool batchable = !L.is_slab() && !V.is_slab() && n > 1;
    for (int i = 1; i < n; ++i) batchable &= dyes[i]->s.iters == L.iters && dyes[i]->s.fused == L.fused && !dyes[i]->s.is_slab();
    if (!batchable) {  // slab-decomposed or differently configured dyes: one step each (same results)
        for (int i = 0; i < n; ++i) TRY(fruid_density_step_dev(dyes[i], vel, dt, diff));
        return 0;
    }
    CU(cudaSetDevice(L.device));
    std::vector<FieldSet*> D;
    std::vector<uint64_t> key{V.uid};
    for (int i = 0; i < n; ++i) {
        FieldSet& s = dyes[i]->s;
        if (V.g.w != s.g.w || V.g.gh != s.g.gh)
            return fail(FRUID_ERR_SIZE_MISMATCH, "density %dx%d vs velocity %dx%d", s.g.w, s.g.gh, V.g.w, V.g.gh);
        if (s.device != V.device) return fail(FRUID_ERR_BAD_ARG, "density and velocity live on different devices");
        TRY(s.consume_uploads());
        D.push_back(&s);
        key.push_back(s.uid);
    }
    // Everything is driven from the first dye's stream: it waits for whatever is still queued on the other dyes'
    // streams (uploads, injects, fills), runs the diffuses (beside the velocity step still in flight: only the
    // advect needs u, v), waits for the velocity step, advects ...
    CU(cudaEventRecord(V.ev, V.stream));
    for (int i = 1; i < n; ++i) {
        CU(cudaEventRecord(D[i]->ev, D[i]->stream));
        CU(stream_wait(L.stream, D[i]->ev));
    }
    if (!g_dens_overlap) {  // one captured graph: fork, diffuses, join, advect -- after the velocity step
        CU(stream_wait(L.stream, V.ev));
        TRY(run_graphed(L.batch_graph, L, D, dt, diff, V.buf[0], V.buf[1], key, [&]() {
            TRY(dens_diffuse_batch_body(D, dt, diff));
            return dens_advect_batch(D, V, dt);
        }));
    } else {
        TRY(run_graphed(L.batch_graph, L, D, dt, diff, nullptr, nullptr, key, [&]() { return dens_diffuse_batch_body(D, dt, diff); }));
        CU(stream_wait(L.stream, V.ev));
        TRY(dens_advect_batch(D, V, dt));
    }
    // ... and the other dyes' streams and the velocity stream continue after it.
    CU(cudaEventRecord(L.ev, L.stream));
    for (int i = 1; i < n; ++i) CU(stream_wait(D[i]->stream, L.ev));
    CU(stream_wait(V.stream, L.ev));
    return 0;
}
extern "C" int fruid_density_step_host(fruid_density_t* d, const float* u, const float* v, float dt, float diff) {
    NEED(d);
    if (!u || !v) return fail(FRUID_ERR_BAD_ARG, "null velocity pointer");
    FieldSet& s = d->s;
    {   // (u, v) = the current host mirrors of a live FluidSim (the c.step(sim.uv(), ..) pattern, main.rs:115)?
        fruid_fluid* vel = nullptr;
        {
            std::lock_guard<std::mutex> lk(g_fluids_mu);
            for (fruid_fluid* f : g_fluids) {
                const FieldSet& fs = f->s;
                if (fs.mirror[FRUID_U] == u && fs.mirror[FRUID_V] == v && fs.mirror_epoch[FRUID_U] == fs.dev_epoch &&
                    fs.mirror_epoch[FRUID_V] == fs.dev_epoch && !fs.is_slab() && fs.device == s.device && fs.g.w == s.g.w &&
                    fs.g.gh == s.g.gh) { vel = f; break; }
            }
        }
        if (vel) return fruid_density_step_dev(d, vel, dt, diff);  // device-resident velocity: nothing to upload
    }
    g_velocity_uploads.fetch_add(1);
    if (s.is_slab()) return fail(FRUID_ERR_BAD_ARG, "slab-decomposed density needs a device-resident velocity (fruid_density_step_dev)");
    // staging fields for the uploaded velocity, kept with the handle (two stream-ordered allocations per call would
    // be handed back to the OS at every synchronisation)
    const size_t bytes = s.g.pitch * (size_t)s.g.h * sizeof(float);
    for (int k = 0; k < 2; ++k)
        if (!s.vel_stage[k]) {
            CU(cudaMalloc(&s.vel_stage[k], bytes));
            CU(cudaMemsetAsync(s.vel_stage[k], 0, bytes, s.stream));  // the row padding stays zero
        }
    float *du = s.vel_stage[0], *dv = s.vel_stage[1];
    const size_t rb = (size_t)s.g.w * sizeof(float);
    CU(cudaMemcpy2DAsync(du, s.g.pitch * sizeof(float), u, rb, rb, (size_t)s.g.h, cudaMemcpyHostToDevice, s.stream));
    CU(cudaMemcpy2DAsync(dv, s.g.pitch * sizeof(float), v, rb, rb, (size_t)s.g.h, cudaMemcpyHostToDevice, s.stream));
    TRY(op_dens_step(s.buf[0], s.buf[1], du, dv, s.buf[2], s.buf[3], s.g, diff, dt, s.iters, s.fused, s.stream));
    CU(cudaStreamSynchronize(s.stream));  // the caller's (possibly pageable) arrays may be reused right away
    return 0;
}
extern "C" void* fruid_density_stream(fruid_density_t* d) { return d ? (void*)d->s.stream : nullptr; }
extern "C" int fruid_density_sync(fruid_density_t* d) { NEED(d); CU(cudaStreamSynchronize(d->s.stream)); TRY(d->s.sync_copies()); CU(cudaGetLastError()); return d->s.check_peer_error(); }

// draw_density (src/main.rs:146-183) on the device: vertices (24 floats per cell, reference push order)
// or compact RGB (3 floats per cell, x + y*w order) of the four dye simulations, copied to the host.
static int draw_density_impl(fruid_density_t* c, fruid_density_t* m, fruid_density_t* y, fruid_density_t* k, float z,
                             float* host, bool verts) {
    if (!c || !m || !y || !k || !host) return fail(FRUID_ERR_BAD_ARG, "null handle or host pointer");
    FieldSet& S = c->s;
    CU(cudaSetDevice(S.device));
    fruid_density_t* all[4] = {c, m, y, k};
    for (fruid_density_t* d : all) {
        if (d->s.g.w != S.g.w || d->s.g.gh != S.g.gh) return fail(FRUID_ERR_SIZE_MISMATCH, "dye fields differ in size");
        if (d->s.is_slab() || d->s.device != S.device) return fail(FRUID_ERR_BAD_ARG, "draw_density needs whole single-GPU fields on one device");
    }
    for (int i = 1; i < 4; ++i) {  // the other dyes' steps must have finished before we read them
        CU(cudaEventRecord(all[i]->s.ev, all[i]->s.stream));
        CU(stream_wait(S.stream, all[i]->s.ev));
    }
    const size_t per_cell = verts ? 24 : 3;
    const size_t bytes = (size_t)S.g.w * S.g.h * per_cell * sizeof(float);
    float* dev = nullptr;
    TRY(S.draw_buffer(bytes, &dev));
    const dim3 grid((S.g.w + 31) / 32, (S.g.h + 31) / 32), block(32, 8);
    const float cw = 2.0f / (float)S.g.w, ch = 2.0f / (float)S.g.h;  // main.rs:147-148
    if (verts)
        LAUNCH("k_draw_density", S.stream, (k_draw_density<true><<<grid, block, 0, S.stream>>>(c->s.buf[0], m->s.buf[0], y->s.buf[0], k->s.buf[0], S.g, z, cw, ch, dev)));
    else
        LAUNCH("k_draw_density_rgb", S.stream, (k_draw_density<false><<<grid, block, 0, S.stream>>>(c->s.buf[0], m->s.buf[0], y->s.buf[0], k->s.buf[0], S.g, z, cw, ch, dev)));
    CU(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, S.stream));
    CU(cudaStreamSynchronize(S.stream));
    return 0;
}
extern "C" int fruid_draw_density(fruid_density_t* c, fruid_density_t* m, fruid_density_t* y, fruid_density_t* k, float z,
                                  float* host_vertices) {
    return draw_density_impl(c, m, y, k, z, host_vertices, true);
}
extern "C" int fruid_density_colors(fruid_density_t* c, fruid_density_t* m, fruid_density_t* y, fruid_density_t* k,
                                    float* host_rgb) {
    return draw_density_impl(c, m, y, k, 0.0f, host_rgb, false);
}

// draw_velocity_lines (src/main.rs:185-216) of sim.uv(): 12 floats per cell {tail, tip} x {pos[3], color[3]}.
extern "C" int fruid_draw_velocity_lines(fruid_fluid_t* f, float z, float* host_vertices) {
    NEED(f);
    if (!host_vertices) return fail(FRUID_ERR_BAD_ARG, "null host pointer");
    FieldSet& S = f->s;
    if (S.is_slab()) return fail(FRUID_ERR_BAD_ARG, "draw_velocity_lines needs a whole single-GPU field");
    const size_t bytes = (size_t)S.g.w * S.g.h * 12 * sizeof(float);
    float* dev = nullptr;
    TRY(S.draw_buffer(bytes, &dev));
    const dim3 grid((S.g.w + 31) / 32, (S.g.h + 31) / 32), block(32, 8);
    const float cw = 2.0f / (float)S.g.w, ch = 2.0f / (float)S.g.h;  // main.rs:186-187
    LAUNCH("k_draw_velocity_lines", S.stream, (k_draw_velocity_lines<<<grid, block, 0, S.stream>>>(S.buf[0], S.buf[1], S.g, z, cw, ch, dev)));
    CU(cudaMemcpyAsync(host_vertices, dev, bytes, cudaMemcpyDeviceToHost, S.stream));
    CU(cudaStreamSynchronize(S.stream));
    return 0;
}

// Page-lock caller-owned host arrays (the Vec<f32> inside an Array2D has a fixed length for
// the object's life) so uploads/downloads run at full PCIe rate.
// Tile shape of the fused Jacobi kernel: R rows per warp x NW warps (tile = 128 x R*NW cells).
extern "C" int fruid_set_tile_config(int rows_per_warp, int warps, int default_fused) {
    if (jacobi_tile_set_cfg(rows_per_warp, warps, default_fused) != 0)
        return fail(FRUID_ERR_BAD_ARG, "unsupported tile config R=%d NW=%d", rows_per_warp, warps);
    g_cfg_epoch.fetch_add(1);
    return 0;
}
// Pure host arithmetic (no CUDA call): the tile shape, fused depth and tiling the library would use for a
// w x h single-GPU field.  out = {rows per warp, warps, sweeps of the first launch, tiles in x, tiles in y, launches}.
extern "C" int fruid_debug_tile_plan(size_t w, size_t h, int iters, int fused, int* out) {
    if (!out) return fail(FRUID_ERR_BAD_ARG, "null out pointer");
    if (w < 3 || h < 3 || w > (size_t)1 << 20 || h > (size_t)1 << 20) return fail(FRUID_ERR_BAD_SIZE, "grid %zux%zu has no tiling", w, h);
    if (iters < 2 || (iters & 1)) return fail(FRUID_ERR_ODD_ITERATIONS, "iterations must be even and >= 2, got %d", iters);
    const Geom g = make_geom(w, h, pitch_for(w));
    const int T = resolve_fused(g, iters, fused);
    if (T <= 1) { out[0] = out[1] = 0; out[2] = 1; out[3] = out[4] = 0; out[5] = iters; return 0; }
    const int nblocks = (iters + T - 1) / T;
    const int T0 = (iters + nblocks - 1) / nblocks;  // op_lin_solve spreads the sweeps evenly
    const TileCfg cfg = jacobi_tile_shape(g, T0);
    if (T0 > jacobi_tile_max_T(cfg.R, cfg.NW)) return fail(FRUID_ERR_BAD_ARG, "fused sweeps T=%d not supported by the current tile config", T0);
    const int H = T0 + (T0 & 1), TH = cfg.R * cfg.NW;
    out[0] = cfg.R; out[1] = cfg.NW; out[2] = T0;
    out[3] = tiles_needed(g.w, 128, 128 - 2 * H);
    out[4] = tiles_needed(g.h, TH, TH - 2 * H);
    out[5] = nblocks;
    return 0;
}
extern "C" int fruid_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return fail(FRUID_ERR_BAD_ARG, "null pointer or zero size");
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();  // callers treat pinning as best effort: do not leave a sticky error for the next launch check
        return fail(FRUID_ERR_CUDA, "cudaHostRegister(%p, %zu) failed: %s", p, bytes, cudaGetErrorString(e));
    }
    return 0;
}
extern "C" int fruid_host_unregister(void* p) {
    if (!p) return fail(FRUID_ERR_BAD_ARG, "null pointer");
    const cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(FRUID_ERR_CUDA, "cudaHostUnregister(%p) failed: %s", p, cudaGetErrorString(e));
    }
    return 0;
}

// --------------------------------------------------------------------------------------
// per-operator entry points on host arrays
// --------------------------------------------------------------------------------------
namespace {
struct Scratch {  // a few temporary device fields for one operator call
    Geom g{};
    std::vector<float*> bufs;
    ~Scratch() { for (float* p : bufs) cudaFree(p); }
    int init(size_t w, size_t h) {
        pdl_chain_reset();
        if (w < 2 || h < 2) return fail(FRUID_ERR_BAD_SIZE, "grid must be at least 2x2 (got %zux%zu)", w, h);
        g = make_geom(w, h, pitch_for(w));
        return 0;
    }
    int put(const float* host, float** out) {
        float* d = nullptr;
        const size_t bytes = g.pitch * (size_t)g.h * sizeof(float);
        CU(cudaMalloc(&d, bytes));
        bufs.push_back(d);
        CU(cudaMemset(d, 0, bytes));
        if (host) CU(cudaMemcpy2D(d, g.pitch * sizeof(float), host, (size_t)g.w * sizeof(float), (size_t)g.w * sizeof(float),
                                  (size_t)g.h, cudaMemcpyHostToDevice));
        *out = d;
        return 0;
    }
    int get(float* host, const float* d) {
        CU(cudaDeviceSynchronize());
        CU(cudaMemcpy2D(host, (size_t)g.w * sizeof(float), d, g.pitch * sizeof(float), (size_t)g.w * sizeof(float), (size_t)g.h,
                        cudaMemcpyDeviceToHost));
        return 0;
    }
};
}  // namespace

extern "C" int fruid_op_add_source(float* x, const float* s, size_t w, size_t h, float dt) {
    if (!x || !s) return fail(FRUID_ERR_BAD_ARG, "null pointer");
    Scratch sc; TRY(sc.init(w, h));
    float *dx = nullptr, *ds = nullptr;
    TRY(sc.put(x, &dx)); TRY(sc.put(s, &ds));
    TRY(op_add_source(dx, ds, nullptr, nullptr, sc.g, dt, 0));
    return sc.get(x, dx);
}
extern "C" int fruid_op_set_bnd(int b, float* x, size_t w, size_t h) {
    if (!x || b < 0 || b > 2) return fail(FRUID_ERR_BAD_ARG, "null pointer or bad bounds kind");
    Scratch sc; TRY(sc.init(w, h));
    float* dx = nullptr;
    TRY(sc.put(x, &dx));
    TRY(op_set_bnd(b, dx, sc.g, 0));
    return sc.get(x, dx);
}
extern "C" int fruid_op_lin_solve(int b, float* x, const float* x0, float* scratch, size_t w, size_t h, float a, float c,
                                  int iters, int fused) {
    if (!x || !x0 || !scratch || b < 0 || b > 2) return fail(FRUID_ERR_BAD_ARG, "null pointer or bad bounds kind");
    Scratch sc; TRY(sc.init(w, h));
    float *dx = nullptr, *dx0 = nullptr, *ds = nullptr;
    TRY(sc.put(x, &dx)); TRY(sc.put(x0, &dx0)); TRY(sc.put(scratch, &ds));
    TRY(op_lin_solve(b, dx, dx0, ds, sc.g, a, c, iters, fused, 0));
    TRY(sc.get(x, dx));
    return sc.get(scratch, ds);
}
extern "C" int fruid_op_diffuse(int b, float* x, const float* x0, float* scratch, size_t w, size_t h, float diff, float dt,
                                int iters, int fused) {
    if (!x || !x0 || !scratch || b < 0 || b > 2) return fail(FRUID_ERR_BAD_ARG, "null pointer or bad bounds kind");
    Scratch sc; TRY(sc.init(w, h));
    float *dx = nullptr, *dx0 = nullptr, *ds = nullptr;
    TRY(sc.put(x, &dx)); TRY(sc.put(x0, &dx0)); TRY(sc.put(scratch, &ds));
    TRY(op_diffuse(b, dx, dx0, ds, sc.g, diff, dt, iters, fused, 0));
    TRY(sc.get(x, dx));
    return sc.get(scratch, ds);
}
extern "C" int fruid_op_advect(int b, float* d, const float* d0, const float* u, const float* v, size_t w, size_t h,
                               float dt) {
    if (!d || !d0 || !u || !v || b < 0 || b > 2) return fail(FRUID_ERR_BAD_ARG, "null pointer or bad bounds kind");
    if (d == d0 || d == u || d == v) return fail(FRUID_ERR_BAD_ARG, "advect output must not alias an input (lib.rs:84)");
    Scratch sc; TRY(sc.init(w, h));
    float *dd = nullptr, *dd0 = nullptr, *du = nullptr, *dv = nullptr;
    TRY(sc.put(d, &dd));
    TRY(sc.put(d0, &dd0));
    if (u == d0) du = dd0; else TRY(sc.put(u, &du));
    if (v == d0) dv = dd0; else if (v == u) dv = du; else TRY(sc.put(v, &dv));
    float* dl[1] = {dd};
    const float* d0l[1] = {dd0};
    const int bl[1] = {b};
    TRY(op_advect(1, dl, d0l, bl, du, dv, sc.g, dt, 0));
    return sc.get(d, dd);
}
extern "C" int fruid_op_project(float* u, float* v, float* p, float* div, float* scratch, size_t w, size_t h, int iters,
                                int fused) {
    if (!u || !v || !p || !div || !scratch) return fail(FRUID_ERR_BAD_ARG, "null pointer");
    Scratch sc; TRY(sc.init(w, h));
    float *du = nullptr, *dv = nullptr, *dp = nullptr, *dd = nullptr, *ds = nullptr;
    TRY(sc.put(u, &du)); TRY(sc.put(v, &dv)); TRY(sc.put(p, &dp)); TRY(sc.put(div, &dd)); TRY(sc.put(scratch, &ds));
    TRY(op_project(du, dv, dp, dd, ds, sc.g, iters, fused, 0));
    TRY(sc.get(u, du)); TRY(sc.get(v, dv)); TRY(sc.get(p, dp)); TRY(sc.get(div, dd));
    return sc.get(scratch, ds);
}

// --------------------------------------------------------------------------------------
// device-pointer variants
// --------------------------------------------------------------------------------------
static int check_dev_layout(const void* p, size_t w, size_t h, size_t pitch) {
    pdl_chain_reset();
    if (w < 2 || h < 2) return fail(FRUID_ERR_BAD_SIZE, "grid must be at least 2x2");
    if (pitch < w || (pitch & 3) || ((uintptr_t)p & 15)) return fail(FRUID_ERR_BAD_ARG, "pitch must be >= w and a multiple of 4 floats; base 16-byte aligned");
    return 0;
}
extern "C" int fruid_dev_lin_solve(int b, float* x, const float* x0, float* scratch, size_t w, size_t h, size_t pitch, float a,
                                   float c, int iters, int fused, void* stream) {
    TRY(check_dev_layout(x, w, h, pitch)); TRY(check_dev_layout(x0, w, h, pitch)); TRY(check_dev_layout(scratch, w, h, pitch));
    float* const x_in = x;
    TRY(op_lin_solve(b, x, x0, scratch, make_geom(w, h, pitch), a, c, iters, fused, (cudaStream_t)stream));
    if (x != x_in)  // odd number of launches: the newest iterate sits in the caller's scratch
        CU(cudaMemcpyAsync(x_in, x, pitch * h * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}
extern "C" int fruid_dev_advect(int b, float* d, const float* d0, const float* u, const float* v, size_t w, size_t h,
                                size_t pitch, float dt, void* stream) {
    TRY(check_dev_layout(d, w, h, pitch)); TRY(check_dev_layout(d0, w, h, pitch));
    float* dl[1] = {d};
    const float* d0l[1] = {d0};
    const int bl[1] = {b};
    return op_advect(1, dl, d0l, bl, u, v, make_geom(w, h, pitch), dt, (cudaStream_t)stream);
}
