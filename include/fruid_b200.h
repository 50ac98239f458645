/*
 * fruid_b200.h -- C ABI of the B200-native stable-fluids time step.
 *
 * Drop-in boundary for Masterchef365/fruid's solver (reference src/lib.rs).
 * The reference has NO pre-existing FFI: its boundary is the Rust crate API
 *     pub type Array2D            src/lib.rs:1
 *     DensitySim::{new,step,density,density_mut}      src/lib.rs:216-245
 *     FluidSim::{new,step,uv,uv_mut,width,height}     src/lib.rs:255-294
 * This header is what a Rust shim keeping those signatures binds with
 * `extern "C"` (INTEGRATION.md shows the shim); include/fruid.hpp is the same
 * mirror in C++ and fruid_b200/api.py in Python (ctypes).
 *
 * Conventions
 *   - Fields are f32, w x h cells INCLUDING the 1-cell border (the constructor
 *     arguments of FluidSim::new / DensitySim::new).  Host arrays are dense,
 *     index (x, y) -> x + y*w (idek_basics::Array2D layout).
 *   - Every function returns 0 (FRUID_OK) or a negative fruid_status;
 *     fruid_last_error() gives a thread-local message for the last failure.
 *     The reference API is infallible (misuse panics); a shim panics on != 0.
 *   - A handle is not thread-safe; distinct handles may be used from distinct
 *     threads (library-global state -- caches, tuning knobs -- is mutex-guarded;
 *     the fruid_set_* / fruid_debug_* knobs are process-wide and meant to be set
 *     while no step is in flight).  Work is enqueued on the handle's CUDA stream;
 *     download/sync calls block until the data is on the host.
 *   - There is no CPU fallback: without a CUDA device every call fails with
 *     FRUID_ERR_CUDA.
 */
#ifndef FRUID_B200_H
#define FRUID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fruid_fluid fruid_fluid_t;     /* device twin of FluidSim   (src/lib.rs:247-253) */
typedef struct fruid_density fruid_density_t; /* device twin of DensitySim (src/lib.rs:210-214) */

typedef enum fruid_status {
    FRUID_OK = 0,
    FRUID_ERR_BAD_SIZE = -1,       /* w < 2 or h < 2 (debug_assert at src/lib.rs:13-14) */
    FRUID_ERR_SIZE_MISMATCH = -2,  /* density vs velocity size (unchecked at src/lib.rs:226) */
    FRUID_ERR_ODD_ITERATIONS = -3, /* sweeps per lin_solve must be even and >= 2 */
    FRUID_ERR_CUDA = -4,
    FRUID_ERR_NO_PEER = -5,
    FRUID_ERR_OOM = -6,
    FRUID_ERR_BAD_ARG = -7
} fruid_status;

/* enum Bounds, src/lib.rs:19-25 (#[repr(i32)], same values). */
typedef enum fruid_bounds { FRUID_POSITIVE = 0, FRUID_NEGX = 1, FRUID_NEGY = 2 } fruid_bounds;

/* FluidSim fields, src/lib.rs:247-253.  uv() -> (U, V); uv_mut() -> (U_PREV, V_PREV).
 * After a step U_PREV holds the pressure and V_PREV the divergence of the
 * second projection (src/lib.rs:207 -> 146), exactly as in the reference. */
typedef enum fruid_fluid_field {
    FRUID_U = 0, FRUID_V = 1, FRUID_U_PREV = 2, FRUID_V_PREV = 3
} fruid_fluid_field;

/* DensitySim fields, src/lib.rs:210-214.  density() -> DENS; density_mut() -> DENS_PREV. */
typedef enum fruid_density_field { FRUID_DENS = 0, FRUID_DENS_PREV = 1 } fruid_density_field;

const char *fruid_last_error(void);
/* Number of kernels this library has launched in this process (all handles). */
uint64_t fruid_kernel_launch_count(void);
/* "sm_100a" build tag + version string. */
const char *fruid_build_info(void);

/* Tuning knob: tile shape of the fused Jacobi kernel (rows per warp x warps per CTA; the tile
 * is 128 x rows_per_warp*warps cells) and the default number of fused sweeps (0 = keep).
 * By default the library chooses per launch: 128x140 tiles (7 rows x 20 warps) and 10 sweeps per launch, smaller tiles
 * (128x64, 128x96) and up to 20 sweeps on grids too small to give every SM a tile.  A call pins the
 * shape for every later launch; (0, 0, T) returns to the automatic choice.  Results never change. */
int fruid_set_tile_config(int rows_per_warp, int warps, int default_fused);
/* Diagnostics, pure host arithmetic: what the library would use for a w x h single-GPU field and `iters` sweeps
 * (fused = 0: library choice).  out[6] = {rows per warp, warps, sweeps per launch, tiles in x, tiles in y, launches
 * per lin_solve}; rows per warp = 0 means one plain kernel per sweep. */
int fruid_debug_tile_plan(size_t w, size_t h, int iters, int fused, int *out);

/* Tuning knob: replay FluidSim::step / DensitySim::step as captured CUDA graphs (default 1). */
int fruid_set_graphs_enabled(int on);
/* Tuning knob: let the dependent kernels of a step be scheduled while their predecessor is still running
 * (programmatic dependent launch; each kernel first waits for the predecessor's completion, so results
 * cannot change).  Default on (environment FRUID_B200_LAUNCH_OVERLAP=0 turns it off at load). */
int fruid_set_launch_overlap(int on);

/* Tuning knob: which advect kernel a single-GPU handle uses.  1 = the gathered fields are staged in shared memory
 * (TMA window of 144 x 40 cells per 128 x 32 block; cells whose back-trace leaves the window gather from global
 * memory), 0 = every tap is a global load, -1 (default) = by grid size (shared-memory window from ~1.5 M cells up;
 * environment FRUID_B200_ADVECT_TILE sets it at load).  Results never change. */
int fruid_set_advect_tile(int mode);

/* Debug knobs of the ordering tests (tests/test_ordering_gpu.py).  fruid_debug_set_schedule: bit 0 = density
 * add_source + diffuse beside the velocity step still in flight, bit 1 = u / v diffuse on two streams, bit 2 = the
 * round-1 launch attribution (programmatic attribute on every launch, even straight after a graph launch or an
 * event wait; kept only to reproduce the stale read it caused).  fruid_debug_set_tail_delay: in the FRUID_DEBUG build
 * (libfruid_b200_dbg.so) the lowest warps of every block wait `ns` nanoseconds before they store, which turns a
 * missing dependency into a deterministic mismatch; the release build ignores it and returns 1. */
int fruid_debug_set_schedule(int mask);
int fruid_debug_set_tail_delay(int ns);
/* Diagnostics: out[0] = divisor c seen before, out[1] = the division mode chosen for it (0 IEEE, 4 / 5 verified
 * reciprocal forms); zhzl = the double-word reciprocal used. */
int fruid_debug_recip_mode(float c, int *out, float *zhzl);

/* Page-lock / unlock a caller-owned host array (e.g. the Vec<f32> inside an Array2D). */
int fruid_host_register(void *p, size_t bytes);
int fruid_host_unregister(void *p);

/* Optional per-kernel timing: when enabled every kernel launch is bracketed by CUDA events
 * on its stream; fruid_profile_report() synchronises the device, sums durations by kernel
 * name into out[0..*n) and resets the log. */
typedef struct fruid_kernel_stat { char name[48]; uint64_t launches; double total_ms; } fruid_kernel_stat;
int fruid_profile_enable(int on);
int fruid_profile_report(fruid_kernel_stat *out, size_t cap, size_t *n);

/* ---- FluidSim (src/lib.rs:255-294) ------------------------------------- */
/* FluidSim::new(width, height), src/lib.rs:256: five zeroed w x h fields. */
int fruid_fluid_create(size_t w, size_t h, fruid_fluid_t **out);
int fruid_fluid_destroy(fruid_fluid_t *f);
/* FluidSim::width / height, src/lib.rs:287-293. */
size_t fruid_fluid_width(const fruid_fluid_t *f);
size_t fruid_fluid_height(const fruid_fluid_t *f);
/* Host <-> device copies of one field (dense w*h floats). */
int fruid_fluid_upload(fruid_fluid_t *f, int field, const float *host);
int fruid_fluid_download(fruid_fluid_t *f, int field, float *host);
/* Asynchronous transfers for page-locked host arrays (fruid_host_register): upload_async stages the
 * field on a copy stream -- it takes effect at the start of the NEXT step() and overlaps the kernels
 * already enqueued; download_async snapshots the field in stream order and copies it out while later
 * steps run.  The host array must stay untouched (upload) / is valid (download) after fruid_*_sync(). */
int fruid_fluid_upload_async(fruid_fluid_t *f, int field, const float *host);
int fruid_fluid_download_async(fruid_fluid_t *f, int field, float *host);
/* `u[pos] = val` through uv_mut(), src/main.rs:98-102, without a full upload. */
int fruid_fluid_inject(fruid_fluid_t *f, int field, size_t x, size_t y, float val);
/* `field.data_mut().fill(val)` without a full upload (cf. fruid_density_fill). */
int fruid_fluid_fill(fruid_fluid_t *f, int field, float val);
/* Sweeps per lin_solve; the reference hard-codes 20 (src/lib.rs:52).  Even only. */
int fruid_fluid_set_iterations(fruid_fluid_t *f, int k);
/* Jacobi sweeps fused per kernel launch (0 = library default; 1 = one sweep per launch). */
int fruid_fluid_set_fused_sweeps(fruid_fluid_t *f, int t);
/* FluidSim::step(dt, visc), src/lib.rs:267-277 -> vel_step src/lib.rs:181-208.  Asynchronous. */
int fruid_fluid_step(fruid_fluid_t *f, float dt, float visc);
/* n identical steps back to back (no host round trip). */
int fruid_fluid_step_n(fruid_fluid_t *f, float dt, float visc, int n);
int fruid_fluid_sync(fruid_fluid_t *f);
/* The handle's cudaStream_t (for event timing / interop). */
void *fruid_fluid_stream(fruid_fluid_t *f);

/* ---- DensitySim (src/lib.rs:216-245) ------------------------------------ */
int fruid_density_create(size_t w, size_t h, fruid_density_t **out);
int fruid_density_destroy(fruid_density_t *d);
int fruid_density_upload(fruid_density_t *d, int field, const float *host);
int fruid_density_download(fruid_density_t *d, int field, float *host);
int fruid_density_upload_async(fruid_density_t *d, int field, const float *host);
int fruid_density_download_async(fruid_density_t *d, int field, float *host);
int fruid_density_inject(fruid_density_t *d, int field, size_t x, size_t y, float val);
/* `density_mut().data_mut().fill(v)`, src/main.rs:105-108. */
int fruid_density_fill(fruid_density_t *d, int field, float val);
int fruid_density_set_iterations(fruid_density_t *d, int k);
int fruid_density_set_fused_sweeps(fruid_density_t *d, int t);
/* DensitySim::step((u, v), dt, diff), src/lib.rs:226-236 -> dens_step src/lib.rs:171-179,
 * with (u, v) = vel's device-resident uv() (the `c.step(sim.uv(), ..)` pattern, src/main.rs:115). */
int fruid_density_step_dev(fruid_density_t *d, fruid_fluid_t *vel, float dt, float diff);
/* The caller pattern of src/main.rs:114-118 -- several DensitySims stepped with the same sim.uv():
 *     c.step(sim.uv(), dt, diff); m.step(sim.uv(), dt, diff); y.step(...); k.step(...);
 * as one call.  Results are those of n fruid_density_step_dev calls, bit for bit; the dyes'
 * diffuses run concurrently and one advect launch moves up to four dyes with a shared back-trace. */
int fruid_density_step_dev_batch(fruid_density_t *const *dyes, int n, fruid_fluid_t *vel, float dt, float diff);
/* Same, with (u, v) given as host arrays, the literal signature of lib.rs:226.  If (u, v) are exactly the
 * arrays bound with fruid_fluid_bind_mirror as the U and V mirrors of a live velocity handle, and both were
 * downloaded after that handle's last change (step, upload, inject), the device-resident velocity is used
 * and nothing is uploaded -- so the reference's `c.step(sim.uv(), dt, diff)` costs no PCIe traffic through a
 * shim that knows nothing about the pairing.  Otherwise the arrays are uploaded first. */
int fruid_density_step_host(fruid_density_t *d, const float *u, const float *v, float dt, float diff);
/* Declares `host` (w*h floats, caller-owned, alive until rebound or the handle is destroyed) to be the
 * caller's mirror of `field`: the array the crate-API shim hands out from uv() / uv_mut().  Its contents must
 * change only through fruid_fluid_download.  See fruid_density_step_host. */
int fruid_fluid_bind_mirror(fruid_fluid_t *f, int field, const float *host);
/* How many times fruid_density_step_host had to upload a velocity pair (diagnostics / tests). */
uint64_t fruid_host_velocity_uploads(void);
int fruid_density_sync(fruid_density_t *d);
void *fruid_density_stream(fruid_density_t *d);

/* ---- Next to the solver: draw_density of the reference viewer, src/main.rs:146-183 ----
 * Computes on the device, from the density() fields of the four dye simulations (c, m, y, k), what
 * the viewer rebuilds on the CPU every frame: the vertex buffer {pos[3], color[3]} x 4 per cell in the
 * reference's push order (i outer, j inner; tl, tr, bl, br) -- host_vertices holds w*h*24 floats -- or
 * just the RGB colour of every cell (host_rgb: w*h*3 floats, cell (x, y) at 3*(x + y*w)). */
int fruid_draw_density(fruid_density_t *c, fruid_density_t *m, fruid_density_t *y, fruid_density_t *k, float z,
                       float *host_vertices);
int fruid_density_colors(fruid_density_t *c, fruid_density_t *m, fruid_density_t *y, fruid_density_t *k,
                         float *host_rgb);
/* draw_velocity_lines(sim.uv(), z), src/main.rs:185-216 (the reference calls it once at
 * init, main.rs:56): per cell, i outer / j inner, the line {tail, tip} as two vertices {pos[3], color[3]}
 * -- host_vertices holds w*h*12 floats.  Bit-exact (+, *, /, sqrt only). */
int fruid_draw_velocity_lines(fruid_fluid_t *f, float z, float *host_vertices);

/* ---- Multi-GPU y-slabs: one process per GPU -----------------------------------
 * Rank `rank` of `world` (<= 8) owns global rows [y0, y1) of the w x gh field plus 22 halo rows
 * of each neighbour.  After creation every rank exports a 64-byte CUDA IPC handle of its device
 * allocation, the ranks exchange them out of band (the Python host side uses torch.distributed),
 * and each rank attaches all peers.  step()/step_dev() then run the same operator sequence as
 * on one GPU with halo rows pulled over NVLink and advection gathers served from the owner's
 * memory; results are bit-identical to the single-GPU run.  upload/download move the OWNED rows
 * (dense w floats per row); inject takes global coordinates.  Every rank must issue the same
 * sequence of step calls.  Teardown is collective as well: a rank may destroy a slab handle only
 * after ALL ranks have finished their work on it (fruid_*_sync on every rank, then a barrier of
 * the host-side process group) -- until then a neighbour may still be pulling halo rows or
 * gathering advection taps from the memory being freed. */
int fruid_fluid_create_slab(size_t w, size_t gh, int rank, int world, fruid_fluid_t **out);
int fruid_fluid_ipc_export(fruid_fluid_t *f, void *handle64);
int fruid_fluid_ipc_attach(fruid_fluid_t *f, int peer_rank, const void *handle64);
int fruid_fluid_slab_rows(const fruid_fluid_t *f, size_t *y0, size_t *y1);
/* Diagnostics (latency of the slab primitives): n halo exchanges of nfields fields / n barriers. */
int fruid_fluid_debug_exchange(fruid_fluid_t *f, int nfields, int n);
int fruid_fluid_debug_barrier(fruid_fluid_t *f, int n);
int fruid_density_create_slab(size_t w, size_t gh, int rank, int world, fruid_density_t **out);
int fruid_density_ipc_export(fruid_density_t *d, void *handle64);
int fruid_density_ipc_attach(fruid_density_t *d, int peer_rank, const void *handle64);

/* ---- Per-operator entry points (parity tests / profiling) ---------------
 * Host arrays in, host arrays out; each call uploads, runs the same kernels the
 * step uses, and downloads.  `fused` = Jacobi sweeps per launch (0 = default).
 * They restate the private fns of src/lib.rs: add_source :5, set_bnd :27,
 * lin_solve :46, diffuse :73, advect :84, project :146, dens_step :171,
 * vel_step :181. */
int fruid_op_add_source(float *x, const float *s, size_t w, size_t h, float dt);
int fruid_op_set_bnd(int b, float *x, size_t w, size_t h);
/* On return x holds the newest iterate (iters must be even); scratch is unspecified
 * (the reference never exposes it: src/lib.rs:213,252). */
int fruid_op_lin_solve(int b, float *x, const float *x0, float *scratch, size_t w, size_t h, float a,
                       float c, int iters, int fused);
int fruid_op_diffuse(int b, float *x, const float *x0, float *scratch, size_t w, size_t h,
                     float diff, float dt, int iters, int fused);
/* d0 may alias u or v (src/lib.rs:204-205); d must be distinct. */
int fruid_op_advect(int b, float *d, const float *d0, const float *u, const float *v, size_t w,
                    size_t h, float dt);
int fruid_op_project(float *u, float *v, float *p, float *div, float *scratch, size_t w, size_t h,
                     int iters, int fused);

/* ---- Device-pointer variants (buffers already in HBM; used by bench/ncu) --
 * All pointers are device pointers to rows of `pitch` floats (pitch >= w,
 * pitch % 4 == 0, base 16-byte aligned); stream is a cudaStream_t (0 = default). */
int fruid_dev_lin_solve(int b, float *x, const float *x0, float *scratch, size_t w, size_t h,
                        size_t pitch, float a, float c, int iters, int fused, void *stream);
int fruid_dev_advect(int b, float *d, const float *d0, const float *u, const float *v, size_t w,
                     size_t h, size_t pitch, float dt, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FRUID_B200_H */
