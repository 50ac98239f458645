// common.cuh -- arithmetic and boundary helpers shared by every kernel.
//
// Bit-exactness contract: the reference (Rust, src/lib.rs) uses only
// IEEE-754 binary32 + - * / with no FMA contraction and no reassociation.  All
// arithmetic below goes through the round-to-nearest intrinsics, which nvcc never
// contracts or reorders; the library is additionally built with -fmad=false and
// without --use_fast_math (default -prec-div=true -ftz=false).
#pragma once
#include <utility>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace fruid {

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

enum : int { B_POSITIVE = 0, B_NEGX = 1, B_NEGY = 2 };  // enum Bounds, src/lib.rs:19-25

// Grid geometry.  The global field is w x gh cells including the border (the constructor
// arguments of FluidSim::new); this device holds `h` consecutive rows of it starting at global
// row gy0 (a y-slab: owned rows plus halo rows of the neighbours).  Single GPU: gy0 = 0, h = gh.
// Rows are addressed LOCALLY (row ly at ly*pitch); boundary logic uses the GLOBAL row ly + gy0.
struct Geom {
    int w, h;       // columns; LOCAL rows held by this device
    int nx, ny;     // GLOBAL interior size (w-2, gh-2), src/lib.rs:12-17
    size_t pitch;   // floats per device row (multiple of 32 for library-owned fields)
    int gy0;        // global row of local row 0
    int gh;         // global number of rows
};

__host__ __device__ __forceinline__ Geom make_geom(size_t w, size_t h, size_t pitch) {
    Geom g;
    g.w = (int)w; g.h = (int)h; g.nx = (int)w - 2; g.ny = (int)h - 2; g.pitch = pitch;
    g.gy0 = 0; g.gh = (int)h;
    return g;
}
__host__ __device__ __forceinline__ Geom make_slab_geom(size_t w, size_t gh, size_t pitch, int gy0, int h_local) {
    Geom g;
    g.w = (int)w; g.h = h_local; g.nx = (int)w - 2; g.ny = (int)gh - 2; g.pitch = pitch;
    g.gy0 = gy0; g.gh = (int)gh;
    return g;
}

__device__ __forceinline__ size_t at(const Geom& g, int x, int y) { return (size_t)x + (size_t)y * g.pitch; }

__device__ __forceinline__ float sgn(bool neg, float v) { return neg ? -v : v; }

// set_bnd (src/lib.rs:27-44) folded into the producer: the thread that owns interior
// cell (i, j) with new value v also writes every border cell that mirrors it.
//   x[0,j]   = sx*x[1,j]      x[nx+1,j] = sx*x[nx,j]      (sx = -1 iff NegX)   lib.rs:31-32
//   x[i,0]   = sy*x[i,1]      x[i,ny+1] = sy*x[i,ny]      (sy = -1 iff NegY)   lib.rs:36-37
//   corner   = 0.5*(row-edge + col-edge) = 0.5*(sy*v + sx*v), v = diagonal interior cell
//                                                                              lib.rs:40-43
// Valid for nx >= 1 and ny >= 1 (for nx == 0 or ny == 0 set_bnd has sequential
// dependencies; those sizes take the serial kernel in serial.cuh).
__device__ __forceinline__ void store_with_bnd(float* __restrict__ out, const Geom& g, int b, int i, int j,
                                               float v) {
    out[at(g, i, j)] = v;
    const bool nxg = (b == B_NEGX), nyg = (b == B_NEGY);
    const bool xl = (i == 1), xr = (i == g.nx), yt = (j == 1), yb = (j == g.ny);
    if (xl | xr | yt | yb) {
        const float vx = sgn(nxg, v), vy = sgn(nyg, v);
        if (xl) out[at(g, 0, j)] = vx;
        if (xr) out[at(g, g.nx + 1, j)] = vx;
        if (yt) out[at(g, i, 0)] = vy;
        if (yb) out[at(g, i, g.ny + 1)] = vy;
        const float c = fmul(0.5f, fadd(vy, vx));
        if (xl && yt) out[at(g, 0, 0)] = c;
        if (xl && yb) out[at(g, 0, g.ny + 1)] = c;
        if (xr && yt) out[at(g, g.nx + 1, 0)] = c;
        if (xr && yb) out[at(g, g.nx + 1, g.ny + 1)] = c;
    }
}

// mix(), src/lib.rs:80-82: (1 - t)*a + t*b, unfused.
__device__ __forceinline__ float mixf(float a, float b, float t) {
    return fadd(fmul(fsub(1.0f, t), a), fmul(t, b));
}

// Programmatic dependent launch: a kernel launched through launch_pdl() may be scheduled while its
// predecessor in the stream is still running, which hides the ~3 us launch latency between the dependent
// kernels of a step (small grids are bound by exactly that).  Every such kernel starts with pdl_enter():
// wait until the predecessor has completed and its writes are visible (nothing is read or written before),
// then allow the successor to be scheduled.  Without the launch attribute both instructions are no-ops.
// The early trigger is taken only while the host-side switch is on (mirrored into this device flag by
// launch_pdl): with the switch off the kernels must behave exactly like plain stream-ordered launches.
static __device__ int g_pdl_trigger_dev = 0;
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (g_pdl_trigger_dev) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_enter_arg(int trigger) {  // same, the switch travels as a kernel argument
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// FRUID_DEBUG build (libfruid_b200_dbg.so, used by the ordering tests only): the lowest warps of every block
// stall for g_dbg_tail_delay_ns before they store their results.  A consumer that starts before its producer
// has COMPLETED (a missing stream / event / programmatic dependency) then reads stale or poisoned rows
// deterministically instead of once in a few hundred runs.
#ifdef FRUID_DEBUG
static __device__ int g_dbg_tail_delay_ns = 0;
__device__ __forceinline__ void dbg_tail_delay_ns(int warp_in_block, int ns) {
    if (warp_in_block >= 3 || ns <= 0) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < (unsigned long long)ns);
}
__device__ __forceinline__ void dbg_tail_delay(int warp_in_block) {
    if (warp_in_block < 3) dbg_tail_delay_ns(warp_in_block, *(volatile int*)&g_dbg_tail_delay_ns);
}
#else
__device__ __forceinline__ void dbg_tail_delay(int) {}
__device__ __forceinline__ void dbg_tail_delay_ns(int, int) {}
#endif
extern int g_dbg_tail_delay_host;  // value of the knob on the host (kernels in other translation units get it as an argument)

// ---- host side of the switch -------------------------------------------------------------------------
// One copy per process (defined in fruid_abi.cu); the kernels' translation units only see the declarations.
struct PdlState {
    bool enabled = true;     // fruid_set_launch_overlap / FRUID_B200_LAUNCH_OVERLAP
    bool unsafe = false;     // FRUID_B200_PDL_UNCHAINED=1: round-1 behaviour (attribute on every launch), experiments only
};
extern PdlState g_pdl;
// Programmatic launch CHAINS.  The attribute is only put on a kernel whose immediate predecessor on the same
// stream is a kernel this thread launched in the same API call: after anything else -- a cudaGraphLaunch, an
// event wait, a copy, the previous API call's work -- the launch is a plain, fully stream-ordered one.  (The
// round-1 library attributed every launch.  Its opt-in density/velocity overlap launched a PDL advect straight
// after a replayed graph and an event wait and showed an intermittent stale read)
bool pdl_chain_continue(cudaStream_t st);  // true: the previous operation on st was a kernel of this chain; marks st open
void pdl_chain_break(cudaStream_t st);     // a non-kernel operation was enqueued on st
void pdl_chain_reset();                    // a new API call begins: nothing is chained across calls
int pdl_dev_flag_sync(int* dev_state, const void* symbol);  // mirrors g_pdl.enabled into a __device__ flag

template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
    static int dev_state = -1;  // per translation unit, like g_pdl_trigger_dev
    if (dev_state != (int)g_pdl.enabled) pdl_dev_flag_sync(&dev_state, (const void*)&g_pdl_trigger_dev);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    const bool chained = pdl_chain_continue(st);
    cfg.numAttrs = (pdl && g_pdl.enabled && (chained || g_pdl.unsafe)) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace fruid
