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

#include "common.cuh"
#include "kernels_basic.cuh"
#include "kernels_serial.cuh"
#include "kernels_vec.cuh"
#include "jacobi_tile.cuh"
#include "slab_comm.cuh"
#include "draw.cuh"

using namespace fruid;

// --------------------------------------------------------------------------------------
// errors, launch accounting
// --------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(e_ == cudaErrorMemoryAllocation ? FRUID_ERR_OOM : FRUID_ERR_CUDA,         \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define TRY(call)              \
    do {                       \
        int r_ = (call);       \
        if (r_ != 0) return r_; \
    } while (0)
// Optional per-kernel timing (bench.py roofline leg): every launch is bracketed by CUDA
// events on the launching stream; fruid_profile_report() synchronises and sums by name.
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
static inline void prof_pre(cudaStream_t st) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r{nullptr, prof_event(), nullptr};
    cudaEventRecord(r.e0, st);
    g_prof.push_back(r);
}
static inline int launched(const char* what, cudaStream_t st) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (g_prof_on) {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (!g_prof.empty() && g_prof.back().name == nullptr) {
            g_prof.back().name = what;
            g_prof.back().e1 = prof_event();
            cudaEventRecord(g_prof.back().e1, st);
        }
    }
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(FRUID_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return 0;
}
#define LAUNCH(name, st, ...)        \
    do {                             \
        prof_pre(st);                \
        __VA_ARGS__;                 \
        TRY(launched(name, st));     \
    } while (0)

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
extern "C" const char* fruid_build_info(void) { return "fruid_b200 0.1 (sm_100a, -fmad=false, IEEE f32 rn)"; }

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
    if (!multi) {
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
static int op_vel_step(float*& u, float*& v, float*& u0, float*& v0, float*& s, float* s2, const Geom& g, float visc,
                       float dt, int iters, int fused, cudaStream_t st) {
    TRY(op_add_source_diffuse(B_NEGX, u0, u, s, s2, g, visc, dt, iters, fused, st));  // lib.rs:190,194
    TRY(op_add_source_diffuse(B_NEGY, v0, v, s, s2, g, visc, dt, iters, fused, st));  // lib.rs:191,197
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

// --------------------------------------------------------------------------------------
// handles
// --------------------------------------------------------------------------------------
// Slab decomposition state of a handle (world == 1: plain single-GPU handle).
struct SlabInfo {
    int rank = 0, world = 1;
    int own0[SLAB_MAX_RANKS + 1] = {0};  // owned global rows of rank r: [own0[r], own0[r+1])
    int row0[SLAB_MAX_RANKS] = {0};      // global row of local row 0 of rank r
    int rows[SLAB_MAX_RANKS] = {0};      // local rows of rank r
    char* peer_base[SLAB_MAX_RANKS] = {nullptr};  // every rank's allocation as mapped here ([rank] = mine)
    int attached = 1;                    // how many of them are mapped (mine counts)
    int ev_pull = 0, ev_bar = 0;         // epochs of the last exchange / barrier
    int ev_adv = 0;                      // epoch of the last advect (src_done / adv_done flags)
    int* adv_block_flag = nullptr;       // per-block "has remote cells" marks of the two-pass advect
    size_t adv_blocks = 0;
};

// A captured CUDA graph of one step for fixed parameters.  The 13-16 launches of a step are
// launch-latency bound on grids up to ~1024^2; replaying a graph removes the per-launch host cost.
static std::atomic<int> g_cfg_epoch{0};  // bumped by every global tuning knob: captured graphs become stale
struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    float dt = 0.f, p = 0.f;       // dt and visc / diff
    int iters = 0, fused = 0;
    const void* a0 = nullptr;      // density: the velocity pointers the graph was captured with
    const void* a1 = nullptr;
    std::vector<const void*> extra;  // batched dye step: the participating handles
    uint64_t kernels = 0;          // kernel launches inside the graph
    int cfg_epoch = -1;            // value of g_cfg_epoch the parameters were recorded under
    bool failed = false;           // capture did not work for these parameters: stay on direct launches
    bool matches(float dt_, float p_, int it, int fu, const void* b0, const void* b1) const {
        return dt == dt_ && p == p_ && iters == it && fused == fu && a0 == b0 && a1 == b1 && cfg_epoch == g_cfg_epoch.load();
    }
    void reset() {
        if (exec) cudaGraphExecDestroy(exec);
        exec = nullptr; failed = false;
    }
};
static bool g_graphs_enabled = true;

struct FieldSet {
    StepGraph graph;
    StepGraph batch_graph;                       // fruid_density_step_dev_batch led by this handle
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;  // fork / join of the batched step (also inside captures)
    Geom g{};
    int n = 0;       // buffers allocated: user fields, then scratch (lib.rs:213,252), then the add_source temporary
    int n_user = 0;  // fields the crate API exposes
    float* buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    char* base = nullptr;  // the one device allocation: [flag block][field 0][field 1]...
    size_t field_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev = nullptr;
    // Asynchronous transfers (pinned host memory): a copy stream, per-field upload staging buffers
    // and one download snapshot, so that PCIe traffic overlaps the kernels of the neighbouring steps.
    cudaStream_t copy_stream = nullptr;
    float* up_stage[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t up_ev[4] = {nullptr, nullptr, nullptr, nullptr};      // staging buffer filled (copy stream)
    cudaEvent_t up_used[4] = {nullptr, nullptr, nullptr, nullptr};    // staging buffer consumed (compute stream)
    bool up_pending[4] = {false, false, false, false};
    float* dl_stage = nullptr;
    cudaEvent_t dl_snap = nullptr, dl_done = nullptr;                 // snapshot taken (compute) / downloaded (copy)
    int iters = 20;  // src/lib.rs:52
    int fused = 0;   // 0 = library default
    int device = 0;
    SlabInfo slab;
    // Slab mode: events of other handles' streams (density steps) that still READ my u, v; awaited
    // lazily right before the step first overwrites u or v instead of at the step's start.
    std::vector<cudaEvent_t> readers;
    int wait_readers() {
        for (cudaEvent_t e : readers) CU(cudaStreamWaitEvent(stream, e, 0));
        readers.clear();
        return 0;
    }

    bool is_slab() const { return slab.world > 1; }
    // local rows [lo, hi) that this rank owns
    int own_lo() const { return slab.own0[slab.rank] - slab.row0[slab.rank]; }
    int own_hi() const { return slab.own0[slab.rank + 1] - slab.row0[slab.rank]; }
    size_t peer_field_bytes(int r) const { return g.pitch * (size_t)slab.rows[r] * sizeof(float); }
    size_t outbox_slot_bytes() const { return g.pitch * (size_t)SLAB_HALO * sizeof(float); }
    // rank r's outbox slot for (event parity, field slot 0..3, direction 0 = towards rank-1, 1 = towards rank+1)
    float* outbox(int r, int parity, int slot, int dir) const {
        return (float*)(slab.peer_base[r] + SLAB_FLAG_BYTES + (size_t)n * peer_field_bytes(r) +
                        (size_t)((parity * 4 + slot) * 2 + dir) * outbox_slot_bytes());
    }
    int phys_index(const float* p) const { return (int)(((const char*)p - (base + SLAB_FLAG_BYTES)) / field_bytes); }
    float* peer_field(int r, int phys) const { return (float*)(slab.peer_base[r] + SLAB_FLAG_BYTES + (size_t)phys * peer_field_bytes(r)); }
    SlabFlags* flags(int r) const { return (SlabFlags*)slab.peer_base[r]; }

    int init(size_t w, size_t gh, int nfields, int rank = 0, int world = 1) {
        if (w < 2 || gh < 2) return fail(FRUID_ERR_BAD_SIZE, "grid must be at least 2x2 (got %zux%zu)", w, gh);
        if (w > (size_t)1 << 20 || gh > (size_t)1 << 20) return fail(FRUID_ERR_BAD_SIZE, "grid too large (%zux%zu)", w, gh);
        if (world < 1 || world > SLAB_MAX_RANKS || rank < 0 || rank >= world)
            return fail(FRUID_ERR_BAD_ARG, "bad rank %d / world %d (at most %d ranks)", rank, world, SLAB_MAX_RANKS);
        CU(cudaGetDevice(&device));
        slab.rank = rank; slab.world = world;
        const int per = (int)gh / world;
        if (world > 1 && (per < 2 * SLAB_HALO || w < 3))
            return fail(FRUID_ERR_BAD_SIZE, "slab decomposition needs at least %d rows per rank (got %d)", 2 * SLAB_HALO, per);
        for (int r = 0; r < world; ++r) slab.own0[r] = r * per;
        slab.own0[world] = (int)gh;
        for (int r = 0; r < world; ++r) {
            const int lo = world > 1 ? std::max(0, slab.own0[r] - SLAB_HALO) : 0;
            const int hi = world > 1 ? std::min((int)gh, slab.own0[r + 1] + SLAB_HALO) : (int)gh;
            slab.row0[r] = lo;
            slab.rows[r] = hi - lo;
        }
        g = make_slab_geom(w, gh, pitch_for(w), slab.row0[rank], slab.rows[rank]);
        n = nfields;
        n_user = nfields - 2;
        CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        field_bytes = g.pitch * (size_t)g.h * sizeof(float);
        // slab mode: outbox = 2 parities x 4 field slots x 2 directions x SLAB_HALO rows (halo exchange staging)
        const size_t total = SLAB_FLAG_BYTES + (size_t)n * field_bytes + (world > 1 ? 16 * outbox_slot_bytes() : 0);
        CU(cudaMalloc(&base, total));
        CU(cudaMemsetAsync(base, 0, total, stream));  // Array2D::new zero-fills, lib.rs:218,257
        for (int i = 0; i < n; ++i) buf[i] = (float*)(base + SLAB_FLAG_BYTES + (size_t)i * field_bytes);
        slab.peer_base[rank] = base;
        CU(cudaStreamSynchronize(stream));
        return 0;
    }
    void release() {
        if (stream) cudaStreamSynchronize(stream);
        graph.reset();
        batch_graph.reset();
        for (int r = 0; r < slab.world; ++r)
            if (r != slab.rank && slab.peer_base[r]) cudaIpcCloseMemHandle(slab.peer_base[r]);
        if (copy_stream) cudaStreamSynchronize(copy_stream);
        for (int i = 0; i < 4; ++i) {
            if (up_stage[i]) cudaFree(up_stage[i]);
            if (up_ev[i]) cudaEventDestroy(up_ev[i]);
            if (up_used[i]) cudaEventDestroy(up_used[i]);
        }
        if (dl_stage) cudaFree(dl_stage);
        if (dl_snap) cudaEventDestroy(dl_snap);
        if (dl_done) cudaEventDestroy(dl_done);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (base) cudaFree(base);
        if (slab.adv_block_flag) cudaFree(slab.adv_block_flag);
        if (ev) cudaEventDestroy(ev);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (stream) { tile_sched_release(stream); cudaStreamDestroy(stream); }
    }
    // Host arrays cover the rows this rank OWNS (the whole field on one GPU), dense w floats per row.
    int upload(int field, const float* host) {
        if (field < 0 || field >= n_user || !host) return fail(FRUID_ERR_BAD_ARG, "bad field %d or null host pointer", field);
        TRY(wait_readers());
        CU(cudaMemcpy2DAsync(buf[field] + (size_t)own_lo() * g.pitch, g.pitch * sizeof(float), host, (size_t)g.w * sizeof(float),
                             (size_t)g.w * sizeof(float), (size_t)(own_hi() - own_lo()), cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));  // host buffer may be pageable and reused by the caller
        return 0;
    }
    int download(int field, float* host) {
        if (field < 0 || field >= n_user || !host) return fail(FRUID_ERR_BAD_ARG, "bad field %d or null host pointer", field);
        CU(cudaMemcpy2DAsync(host, (size_t)g.w * sizeof(float), buf[field] + (size_t)own_lo() * g.pitch, g.pitch * sizeof(float),
                             (size_t)g.w * sizeof(float), (size_t)(own_hi() - own_lo()), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        return check_peer_error();
    }
    int ensure_copy_stream() {
        if (!copy_stream) CU(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        return 0;
    }
    // Enqueue host -> staging on the copy stream; the next step() moves the staged rows into the field.
    // The host array must be page-locked (fruid_host_register) and stay untouched until that step.
    int upload_async(int field, const float* host) {
        if (field < 0 || field >= n_user || !host) return fail(FRUID_ERR_BAD_ARG, "bad field %d or null host pointer", field);
        TRY(ensure_copy_stream());
        const size_t rows = (size_t)(own_hi() - own_lo());
        if (!up_stage[field]) {
            CU(cudaMalloc(&up_stage[field], g.pitch * rows * sizeof(float)));
            CU(cudaEventCreateWithFlags(&up_ev[field], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&up_used[field], cudaEventDisableTiming));
        } else {
            CU(cudaStreamWaitEvent(copy_stream, up_used[field], 0));  // previous staged upload has been consumed
        }
        CU(cudaMemcpy2DAsync(up_stage[field], g.pitch * sizeof(float), host, (size_t)g.w * sizeof(float),
                             (size_t)g.w * sizeof(float), rows, cudaMemcpyHostToDevice, copy_stream));
        CU(cudaEventRecord(up_ev[field], copy_stream));
        up_pending[field] = true;
        return 0;
    }
    // Called at the start of a step: staged uploads become the fields' contents (device-to-device).
    int consume_uploads() {
        for (int f = 0; f < 4; ++f) {
            if (!up_pending[f]) continue;
            TRY(wait_readers());
            CU(cudaStreamWaitEvent(stream, up_ev[f], 0));
            const size_t rows = (size_t)(own_hi() - own_lo());
            CU(cudaMemcpyAsync(buf[f] + (size_t)own_lo() * g.pitch, up_stage[f], g.pitch * rows * sizeof(float),
                               cudaMemcpyDeviceToDevice, stream));
            CU(cudaEventRecord(up_used[f], stream));
            up_pending[f] = false;
        }
        return 0;
    }
    // Snapshot the field on the compute stream, then device -> host on the copy stream (overlaps the
    // next step).  The host array is valid after sync_copies().
    int download_async(int field, float* host) {
        if (field < 0 || field >= n_user || !host) return fail(FRUID_ERR_BAD_ARG, "bad field %d or null host pointer", field);
        TRY(ensure_copy_stream());
        const size_t rows = (size_t)(own_hi() - own_lo());
        if (!dl_stage) {
            CU(cudaMalloc(&dl_stage, g.pitch * rows * sizeof(float)));
            CU(cudaEventCreateWithFlags(&dl_snap, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&dl_done, cudaEventDisableTiming));
        } else {
            CU(cudaStreamWaitEvent(stream, dl_done, 0));  // previous download has left the snapshot buffer
        }
        CU(cudaMemcpyAsync(dl_stage, buf[field] + (size_t)own_lo() * g.pitch, g.pitch * rows * sizeof(float),
                           cudaMemcpyDeviceToDevice, stream));
        CU(cudaEventRecord(dl_snap, stream));
        CU(cudaStreamWaitEvent(copy_stream, dl_snap, 0));
        CU(cudaMemcpy2DAsync(host, (size_t)g.w * sizeof(float), dl_stage, g.pitch * sizeof(float), (size_t)g.w * sizeof(float),
                             rows, cudaMemcpyDeviceToHost, copy_stream));
        CU(cudaEventRecord(dl_done, copy_stream));
        return 0;
    }
    int sync_copies() {
        if (copy_stream) CU(cudaStreamSynchronize(copy_stream));
        return 0;
    }
    // (x, y) are GLOBAL cell coordinates; ranks that do not own row y ignore the write (their
    // halo copy is refreshed from the owner at the start of the next step).
    int inject(int field, size_t x, size_t y, float val) {
        if (field < 0 || field >= n_user) return fail(FRUID_ERR_BAD_ARG, "bad field %d", field);
        if (x >= (size_t)g.w || y >= (size_t)g.gh) return fail(FRUID_ERR_BAD_ARG, "cell (%zu,%zu) outside %dx%d", x, y, g.w, g.gh);
        const int ly = (int)y - g.gy0;
        if (ly < own_lo() || ly >= own_hi()) return 0;
        TRY(wait_readers());
        CU(cudaMemcpyAsync(buf[field] + x + (size_t)ly * g.pitch, &val, sizeof(float), cudaMemcpyHostToDevice, stream));
        return 0;
    }
    int set_iters(int k) {
        if (k < 2 || (k & 1)) return fail(FRUID_ERR_ODD_ITERATIONS, "iterations must be even and >= 2, got %d", k);
        iters = k;
        return 0;
    }
    int set_fused(int t) {
        if (t < 0 || (t > 1 && (t & 1))) return fail(FRUID_ERR_BAD_ARG, "fused sweeps must be 0 (default), 1 or even, got %d", t);
        if (is_slab() && (t == 1 || t > SLAB_MAX_T)) return fail(FRUID_ERR_BAD_ARG, "slab mode needs 2..%d fused sweeps, got %d", SLAB_MAX_T, t);
        fused = t;
        return 0;
    }
    int slab_fused() const {  // fused depth used in slab mode
        int t = resolve_fused(g, iters, fused);
        if (t > SLAB_MAX_T) t = SLAB_MAX_T;
        return t < 2 ? 2 : t;
    }
    int check_peer_error() {
        if (!is_slab()) return 0;
        int err = 0;
        CU(cudaMemcpy(&err, &flags(slab.rank)->error, sizeof(int), cudaMemcpyDeviceToHost));
        if (err) return fail(FRUID_ERR_NO_PEER, "rank %d timed out waiting for a peer (halo exchange / barrier)", slab.rank);
        return 0;
    }
    int need_peers() const {
        if (is_slab() && slab.attached < slab.world)
            return fail(FRUID_ERR_NO_PEER, "slab handle: %d of %d peers attached (fruid_*_ipc_attach)", slab.attached, slab.world);
        return 0;
    }
    int ipc_export(void* out64) {
        if (!out64) return fail(FRUID_ERR_BAD_ARG, "null handle buffer");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, base));
        memcpy(out64, &h, 64);
        return 0;
    }
    int ipc_attach(int peer, const void* in64) {
        if (!in64 || peer < 0 || peer >= slab.world) return fail(FRUID_ERR_BAD_ARG, "bad peer rank %d", peer);
        if (peer == slab.rank || slab.peer_base[peer]) return 0;
        cudaIpcMemHandle_t h;
        memcpy(&h, in64, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(FRUID_ERR_NO_PEER, "cudaIpcOpenMemHandle(rank %d) failed: %s", peer, cudaGetErrorString(e));
        slab.peer_base[peer] = (char*)p;
        slab.attached += 1;
        return 0;
    }
    // Halo exchange of up to 4 fields: publish my boundary rows in my outbox, pull the neighbours'.
    int exchange(std::initializer_list<float*> fields, int wait_adv_event = 0) {
        if (!is_slab()) return 0;
        PullArgs A{};
        A.wait_adv_event = wait_adv_event;
        A.world = slab.world;
        const int me = slab.rank;
        A.me = me; A.pitch = g.pitch; A.event = ++slab.ev_pull;
        A.nbr[0] = me > 0 ? me - 1 : -1;
        A.nbr[1] = me + 1 < slab.world ? me + 1 : -1;
        A.mine = flags(me);
        for (int d = 0; d < 2; ++d) A.peer[d] = A.nbr[d] >= 0 ? flags(A.nbr[d]) : nullptr;
        const int par = A.event & 1;
        int nj = 0, slot = 0;
        for (float* fp : fields) {
            PullJob up{}, dn{};
            if (A.nbr[0] >= 0) {
                // the upper neighbour's bottom halo = my first owned rows; my top halo = its last owned rows,
                // which it publishes in its "towards rank+1" slot
                up.out_rows = slab.row0[me - 1] + slab.rows[me - 1] - slab.own0[me];
                up.mine = fp + (size_t)own_lo() * g.pitch;
                up.outbox = outbox(me, par, slot, 0);
                up.in_rows = slab.own0[me] - slab.row0[me];
                up.dst = fp;
                up.src = outbox(me - 1, par, slot, 1);
            }
            if (A.nbr[1] >= 0) {
                dn.out_rows = slab.own0[me + 1] - slab.row0[me + 1];
                dn.mine = fp + (size_t)(own_hi() - dn.out_rows) * g.pitch;
                dn.outbox = outbox(me, par, slot, 1);
                dn.in_rows = slab.row0[me] + slab.rows[me] - slab.own0[me + 1];
                dn.dst = fp + (size_t)own_hi() * g.pitch;
                dn.src = outbox(me + 1, par, slot, 0);
            }
            A.job[nj++] = up;
            A.job[nj++] = dn;
            ++slot;
        }
        A.njobs = nj;
        // enough CTAs to keep ~2-3 MB in flight over NVLink for wide grids, one pass for narrow ones
        const int ctas = (int)std::min<size_t>(64, std::max<size_t>(8, (g.pitch * SLAB_HALO / 4 + 1023) / 1024));
        LAUNCH("k_slab_pull", stream, (k_slab_pull<<<dim3(ctas, nj), 256, 0, stream>>>(A)));
        return 0;
    }
    int barrier() {
        if (!is_slab()) return 0;
        BarrierArgs A{};
        A.event = ++slab.ev_bar; A.me = slab.rank; A.world = slab.world; A.mine = flags(slab.rank);
        for (int r = 0; r < slab.world; ++r) A.peer[r] = flags(r);
        LAUNCH("k_slab_barrier", stream, (k_slab_barrier<<<1, 32, 0, stream>>>(A)));
        return 0;
    }
    // "I have finished writing the fields the next advect gathers from" (which = 0) / "I have finished
    // that advect" (which = 1): broadcast to every rank, nobody waits here.
    int signal_all(int which) {
        SignalArgs A{};
        A.event = slab.ev_adv; A.me = slab.rank; A.world = slab.world; A.which = which; A.mine = flags(slab.rank);
        for (int r = 0; r < slab.world; ++r) A.peer[r] = flags(r);
        LAUNCH("k_slab_signal_all", stream, (k_slab_signal_all<<<1, 32, 0, stream>>>(A)));
        return 0;
    }
    // Wait until every rank has signalled `which` for advect epoch `event` (before overwriting gathered fields).
    int wait_all(int which, int event) {
        if (!is_slab() || event <= 0) return 0;
        SignalArgs A{};
        A.event = event; A.me = slab.rank; A.world = slab.world; A.which = which; A.mine = flags(slab.rank);
        LAUNCH("k_slab_wait_all", stream, (k_slab_wait_all<<<1, 32, 0, stream>>>(A)));
        return 0;
    }
    // Gather table of the coming advect (epoch slab.ev_adv): where the owners keep field `fp`
    // (f = slot 0/1), as pointers indexed by GLOBAL row; plus the flags of the two-pass scheme.
    int fill_table(SlabTable& t, int f, const float* fp) {
        t.world = slab.world;
        t.me = slab.rank;
        t.my_lo = slab.own0[slab.rank]; t.my_hi = slab.own0[slab.rank + 1];
        for (int r = 0; r <= slab.world; ++r) t.own0[r] = slab.own0[r];
        const int phys = phys_index(fp);
        for (int r = 0; r < slab.world; ++r) t.base[f][r] = peer_field(r, phys) - (ptrdiff_t)slab.row0[r] * (ptrdiff_t)g.pitch;
        const size_t blocks = (size_t)(((g.w + 3) / 4 + 31) / 32) * (size_t)((g.h + 7) / 8);
        if (blocks > slab.adv_blocks) {
            if (slab.adv_block_flag) CU(cudaFree(slab.adv_block_flag));
            CU(cudaMalloc(&slab.adv_block_flag, (2 * blocks + 2) * sizeof(int)));  // flags, pass-2 work list, pass-2 block counter
            CU(cudaMemsetAsync(slab.adv_block_flag, 0, (2 * blocks + 2) * sizeof(int), stream));
            slab.adv_blocks = blocks;
        }
        t.block_flag = slab.adv_block_flag;
        t.block_list = slab.adv_block_flag + slab.adv_blocks;
        t.pass2_done = slab.adv_block_flag + 2 * slab.adv_blocks + 1;
        for (int r = 0; r < slab.world; ++r) t.flags[r] = flags(r);
        t.src_done = flags(slab.rank)->src_done;
        t.error = &flags(slab.rank)->error;
        t.event = slab.ev_adv;
        return 0;
    }
    const float* global_rows(const float* fp) const { return fp - (ptrdiff_t)g.gy0 * (ptrdiff_t)g.pitch; }
};

struct fruid_fluid { FieldSet s; };    // buf: 0=u 1=v 2=u_prev 3=v_prev 4=scratch (lib.rs:247-253) 5=add_source temporary
struct fruid_density { FieldSet s; };  // buf: 0=dens 1=dens_prev 2=scratch (lib.rs:210-214) 3=add_source temporary

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
    CU(cudaStreamWaitEvent(st, vel_ready, 0)); // u, v of this step
    SlabTable tab{};
    TRY(S.fill_table(tab, 0, x0));
    float* d[1] = {x};
    const float* d0[1] = {S.global_rows(x0)};
    const int b[1] = {B_POSITIVE};
    const int ly0 = std::max(S.own_lo(), 1 - g.gy0), ly1 = std::min(S.own_hi(), g.gh - 1 - g.gy0);
    return op_advect(1, d, d0, b, u, v, g, dt, st, &tab, ly0, ly1);                     // lib.rs:178 (pass 2 broadcasts adv_done)
}

// Runs `body` (which enqueues one step on S.stream) either directly or as a replayed CUDA graph.
// A graph is captured on the second step with given parameters (the first warms lazy allocations),
// and only kept if the step left the buffer roles unchanged (an even number of launches per
// lin_solve), because the graph bakes the pointers in.
template <class Body>
static int run_graphed(StepGraph& G, FieldSet& S, const std::vector<FieldSet*>& parts, float dt, float p, const void* a0,
                       const void* a1, const std::vector<const void*>& extra, Body body) {
    bool usable = g_graphs_enabled && !g_prof_on;
    for (FieldSet* P : parts) usable &= !P->is_slab();
    if (!usable) return body();
    if (!G.matches(dt, p, S.iters, S.fused, a0, a1) || G.extra != extra) {
        // first step with these parameters: run it directly (it warms lazy allocations, tensor maps and
        // the divisor check, none of which may happen inside a capture) and remember the parameters
        G.reset();
        G.dt = dt; G.p = p; G.iters = S.iters; G.fused = S.fused; G.a0 = a0; G.a1 = a1; G.extra = extra;
        G.cfg_epoch = g_cfg_epoch.load();
        return body();
    }
    if (G.exec) {
        CU(cudaGraphLaunch(G.exec, S.stream));
        g_launches.fetch_add(G.kernels, std::memory_order_relaxed);
        return 0;
    }
    if (G.failed) return body();
    // second step with the same parameters: capture it
    std::vector<float*> before;
    for (FieldSet* P : parts) before.insert(before.end(), P->buf, P->buf + 6);
    auto restore = [&]() { size_t k = 0; for (FieldSet* P : parts) for (int i = 0; i < 6; ++i) P->buf[i] = before[k++]; };
    auto unchanged = [&]() { size_t k = 0; bool same = true; for (FieldSet* P : parts) for (int i = 0; i < 6; ++i) same &= (P->buf[i] == before[k++]); return same; };
    const uint64_t k0 = g_launches.load();
    if (cudaStreamBeginCapture(S.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        G.failed = true;
        return body();
    }
    const int r = body();
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(S.stream, &graph);
    const uint64_t nk = g_launches.load() - k0;
    g_launches.fetch_sub(nk, std::memory_order_relaxed);  // nothing has run yet
    if (r != 0 || e != cudaSuccess || !graph || !unchanged() ||
        cudaGraphInstantiate(&G.exec, graph, nullptr, nullptr, 0) != cudaSuccess) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        restore();
        G.exec = nullptr;
        G.failed = true;
        return body();  // the captured work was never launched: run the step directly
    }
    cudaGraphDestroy(graph);
    G.kernels = nk;
    CU(cudaGraphLaunch(G.exec, S.stream));
    g_launches.fetch_add(nk, std::memory_order_relaxed);
    return 0;
}
template <class Body>
static int run_step_graphed(FieldSet& S, float dt, float p, const void* a0, const void* a1, Body body) {
    return run_graphed(S.graph, S, {&S}, dt, p, a0, a1, {}, body);
}
// Tuning / debugging knob: replay steps as CUDA graphs (default on).
// Tuning / debugging knob: programmatic dependent launch between the kernels of a step (default on).
namespace { struct PdlEnv { PdlEnv() { if (const char* e = getenv("FRUID_B200_LAUNCH_OVERLAP")) g_pdl_enabled = atoi(e) != 0; } } g_pdl_env; }
extern "C" int fruid_set_launch_overlap(int on) { g_pdl_enabled = on != 0; g_cfg_epoch.fetch_add(1); return 0; }
extern "C" int fruid_set_graphs_enabled(int on) { g_graphs_enabled = on != 0; g_cfg_epoch.fetch_add(1); return 0; }

#define NEED(h) do { if (!(h)) return fail(FRUID_ERR_BAD_ARG, "null handle"); CU(cudaSetDevice((h)->s.device)); } while (0)

extern "C" int fruid_fluid_create(size_t w, size_t h, fruid_fluid_t** out) {
    if (!out) return fail(FRUID_ERR_BAD_ARG, "null out pointer");
    *out = nullptr;
    fruid_fluid* f = new (std::nothrow) fruid_fluid();
    if (!f) return fail(FRUID_ERR_OOM, "host allocation failed");
    int r = f->s.init(w, h, 6);
    if (r) { f->s.release(); delete f; return r; }
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
extern "C" int fruid_fluid_step(fruid_fluid_t* f, float dt, float visc) {
    NEED(f);
    FieldSet& s = f->s;
    TRY(s.consume_uploads());
    if (s.is_slab()) return op_vel_step_slab(s, visc, dt);
    return run_step_graphed(s, dt, visc, nullptr, nullptr, [&]() {
        return op_vel_step(s.buf[0], s.buf[1], s.buf[2], s.buf[3], s.buf[4], s.buf[5], s.g, visc, dt, s.iters, s.fused, s.stream);
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
extern "C" int fruid_density_fill(fruid_density_t* d, int field, float val) {
    NEED(d);
    FieldSet& s = d->s;
    if (field < 0 || field > 1) return fail(FRUID_ERR_BAD_ARG, "bad field %d", field);
    if (val == 0.0f && !std::signbit(val)) {
        CU(cudaMemsetAsync(s.buf[field], 0, s.g.pitch * (size_t)s.g.h * sizeof(float), s.stream));
        return 0;
    }
    LAUNCH("k_fill", s.stream, k_fill<<<cell_grid(s.g.w, s.g.h), cell_block(), 0, s.stream>>>(s.buf[field], s.g, val));
    return 0;
}
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
    // u,v are produced on vel's stream and consumed on ours; order both ways.
    CU(cudaEventRecord(vs.ev, vs.stream));
    if (!s.is_slab()) CU(cudaStreamWaitEvent(s.stream, vs.ev, 0));
    if (s.is_slab())
        TRY(op_dens_step_slab(s, vs.buf[0], vs.buf[1], diff, dt, vs.ev));
    else
        TRY(run_step_graphed(s, dt, diff, vs.buf[0], vs.buf[1], [&]() {
            return op_dens_step(s.buf[0], s.buf[1], vs.buf[0], vs.buf[1], s.buf[2], s.buf[3], s.g, diff, dt, s.iters, s.fused, s.stream);
        }));
    CU(cudaEventRecord(s.ev, s.stream));
    if (s.is_slab()) {
        if (std::find(vs.readers.begin(), vs.readers.end(), s.ev) == vs.readers.end()) vs.readers.push_back(s.ev);
    } else {
        CU(cudaStreamWaitEvent(vs.stream, s.ev, 0));
    }
    return 0;
}
// Several dye fields advanced by one velocity field (the four DensitySims of main.rs:114-118): every
// dye's add_source + diffuse runs on its own stream (concurrently on small grids), then ONE advect
// launch moves up to four dyes with a shared back-trace (u, v are read once instead of once per dye).
// Results are those of n separate fruid_density_step_dev calls, bit for bit.
static int dens_step_batch_body(const std::vector<FieldSet*>& D, FieldSet& V, float dt, float diff) {
    FieldSet& L = *D[0];
    CU(cudaEventRecord(L.ev_fork, L.stream));
    for (size_t i = 0; i < D.size(); ++i) {
        FieldSet& s = *D[i];
        if (i) CU(cudaStreamWaitEvent(s.stream, L.ev_fork, 0));
        TRY(op_add_source_diffuse(B_POSITIVE, s.buf[1], s.buf[0], s.buf[2], s.buf[3], s.g, diff, dt, s.iters, s.fused, s.stream));  // lib.rs:172,175
        if (i) {
            CU(cudaEventRecord(s.ev_join, s.stream));
            CU(cudaStreamWaitEvent(L.stream, s.ev_join, 0));
        }
    }
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
    std::vector<const void*> key;
    for (int i = 0; i < n; ++i) {
        FieldSet& s = dyes[i]->s;
        if (V.g.w != s.g.w || V.g.gh != s.g.gh)
            return fail(FRUID_ERR_SIZE_MISMATCH, "density %dx%d vs velocity %dx%d", s.g.w, s.g.gh, V.g.w, V.g.gh);
        if (s.device != V.device) return fail(FRUID_ERR_BAD_ARG, "density and velocity live on different devices");
        TRY(s.consume_uploads());
        D.push_back(&s);
        key.push_back(&s);
    }
    // Everything is driven from the first dye's stream: it waits for the velocity step and for whatever
    // is still queued on the other dyes' streams (uploads, injects, fills) ...
    CU(cudaEventRecord(V.ev, V.stream));
    CU(cudaStreamWaitEvent(L.stream, V.ev, 0));
    for (int i = 1; i < n; ++i) {
        CU(cudaEventRecord(D[i]->ev, D[i]->stream));
        CU(cudaStreamWaitEvent(L.stream, D[i]->ev, 0));
    }
    TRY(run_graphed(L.batch_graph, L, D, dt, diff, V.buf[0], V.buf[1], key, [&]() { return dens_step_batch_body(D, V, dt, diff); }));
    // ... and the other dyes' streams and the velocity stream continue after it.
    CU(cudaEventRecord(L.ev, L.stream));
    for (int i = 1; i < n; ++i) CU(cudaStreamWaitEvent(D[i]->stream, L.ev, 0));
    CU(cudaStreamWaitEvent(V.stream, L.ev, 0));
    return 0;
}
extern "C" int fruid_density_step_host(fruid_density_t* d, const float* u, const float* v, float dt, float diff) {
    NEED(d);
    if (!u || !v) return fail(FRUID_ERR_BAD_ARG, "null velocity pointer");
    FieldSet& s = d->s;
    if (s.is_slab()) return fail(FRUID_ERR_BAD_ARG, "slab-decomposed density needs a device-resident velocity (fruid_density_step_dev)");
    const size_t bytes = s.g.pitch * (size_t)s.g.h * sizeof(float);
    float *du = nullptr, *dv = nullptr;
    CU(cudaMallocAsync(&du, bytes, s.stream));
    CU(cudaMallocAsync(&dv, bytes, s.stream));
    const size_t rb = (size_t)s.g.w * sizeof(float);
    CU(cudaMemcpy2DAsync(du, s.g.pitch * sizeof(float), u, rb, rb, (size_t)s.g.h, cudaMemcpyHostToDevice, s.stream));
    CU(cudaMemcpy2DAsync(dv, s.g.pitch * sizeof(float), v, rb, rb, (size_t)s.g.h, cudaMemcpyHostToDevice, s.stream));
    int r = op_dens_step(s.buf[0], s.buf[1], du, dv, s.buf[2], s.buf[3], s.g, diff, dt, s.iters, s.fused, s.stream);
    cudaFreeAsync(du, s.stream);
    cudaFreeAsync(dv, s.stream);
    TRY(r);
    CU(cudaStreamSynchronize(s.stream));
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
        CU(cudaStreamWaitEvent(S.stream, all[i]->s.ev, 0));
    }
    const size_t per_cell = verts ? 24 : 3;
    const size_t bytes = (size_t)S.g.w * S.g.h * per_cell * sizeof(float);
    float* dev = nullptr;
    CU(cudaMallocAsync(&dev, bytes, S.stream));
    const dim3 grid((S.g.w + 31) / 32, (S.g.h + 31) / 32), block(32, 8);
    const float cw = 2.0f / (float)S.g.w, ch = 2.0f / (float)S.g.h;  // main.rs:147-148
    if (verts)
        LAUNCH("k_draw_density", S.stream, (k_draw_density<true><<<grid, block, 0, S.stream>>>(c->s.buf[0], m->s.buf[0], y->s.buf[0], k->s.buf[0], S.g, z, cw, ch, dev)));
    else
        LAUNCH("k_draw_density_rgb", S.stream, (k_draw_density<false><<<grid, block, 0, S.stream>>>(c->s.buf[0], m->s.buf[0], y->s.buf[0], k->s.buf[0], S.g, z, cw, ch, dev)));
    CU(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, S.stream));
    CU(cudaFreeAsync(dev, S.stream));
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
    CU(cudaMallocAsync(&dev, bytes, S.stream));
    const dim3 grid((S.g.w + 31) / 32, (S.g.h + 31) / 32), block(32, 8);
    const float cw = 2.0f / (float)S.g.w, ch = 2.0f / (float)S.g.h;  // main.rs:186-187
    LAUNCH("k_draw_velocity_lines", S.stream, (k_draw_velocity_lines<<<grid, block, 0, S.stream>>>(S.buf[0], S.buf[1], S.g, z, cw, ch, dev)));
    CU(cudaMemcpyAsync(host_vertices, dev, bytes, cudaMemcpyDeviceToHost, S.stream));
    CU(cudaFreeAsync(dev, S.stream));
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
extern "C" int fruid_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return fail(FRUID_ERR_BAD_ARG, "null pointer or zero size");
    CU(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    return 0;
}
extern "C" int fruid_host_unregister(void* p) {
    if (!p) return fail(FRUID_ERR_BAD_ARG, "null pointer");
    CU(cudaHostUnregister(p));
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
