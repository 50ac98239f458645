// slab_comm.cuh -- multi-GPU y-slab decomposition over NVLink peer memory.
//
// One process per GPU.  Rank r owns global rows [own0[r], own0[r+1]) and keeps HALO extra rows
// of each neighbour on either side.  All device state of a handle lives in ONE allocation
// (flags + fields) that is exported with cudaIpcGetMemHandle and mapped by every peer, so
//   * halo rows are PULLED from the neighbour's owned rows with plain peer loads, and
//   * advection gathers at any distance read the owner's copy directly (slab_row()).
// Cross-GPU ordering uses monotonically increasing epoch flags in peer memory, written with
// system-scope release and polled with system-scope acquire.  Every rank executes the same
// sequence of events, a wait only ever targets an event its peer signals without waiting for
// anything later, hence no cycles.  Spin loops give up after SPIN_TIMEOUT_NS and raise an error
// word instead of hanging the GPU.
#pragma once
#include "common.cuh"

namespace fruid {

constexpr int SLAB_HALO = 22;       // rows kept of each neighbour: two blocks of 10 fused sweeps + 2
constexpr int SLAB_MAX_T = 10;      // fused sweeps per launch allowed in slab mode
constexpr int SLAB_MAX_RANKS = 8;
constexpr size_t SLAB_FLAG_BYTES = 4096;
constexpr unsigned long long SPIN_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

// Flag block at the start of every rank's allocation (written by peers, read locally).
struct SlabFlags {
    int ready[SLAB_MAX_RANKS];     // ready[s]    >= e : rank s has produced everything up to event e
    int bar[SLAB_MAX_RANKS];       // bar[s]      >= e : rank s has reached barrier e
    int src_done[SLAB_MAX_RANKS];  // src_done[s] >= e : rank s has finished WRITING the fields advect e gathers from
    int adv_done[SLAB_MAX_RANKS];  // adv_done[s] >= e : rank s has finished advect e (stopped READING those fields)
    int error;                     // set when a spin timed out
    int pull_done;                 // CTA counter of the running pull kernel (self-resetting)
};

__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ld_relaxed_sys(const int* p) {
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Spin until *flag >= e (false on timeout): relaxed system-scope polls, one acquire fence at the end.
__device__ __forceinline__ bool spin_until(const int* flag, int e, int* err) {
    if (ld_relaxed_sys(flag) < e) {
        const unsigned long long t0 = globaltimer_ns();
        unsigned it = 0;
        while (ld_relaxed_sys(flag) < e) {
            if ((++it & 1023u) == 0 && globaltimer_ns() - t0 > SPIN_TIMEOUT_NS) {
                *err = 1;
                return false;
            }
        }
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    return true;
}

// One direction of one field in a halo exchange: my boundary rows go to my OUTBOX (a staging area
// in my own allocation that the neighbour reads), the neighbour's outbox rows come into my halo rows.
struct PullJob {
    const float* mine;     // first of my owned rows that the neighbour in this direction needs
    float* outbox;         // my outbox slot for (this event's parity, field, direction)
    const float* src;      // the neighbour's outbox slot holding MY halo rows (peer-mapped pointer)
    float* dst;            // my halo rows
    int out_rows, in_rows; // rows I publish / rows I receive (0 when there is no neighbour)
};
struct PullArgs {
    PullJob job[8];        // up to 4 fields x 2 directions; job 2f = up (rank-1), 2f+1 = down (rank+1)
    int njobs;
    size_t pitch;
    int event;             // epoch of this exchange
    int me;
    int nbr[2];            // neighbour ranks (-1 = none): up, down
    SlabFlags* mine;       // my flag block
    SlabFlags* peer[2];    // neighbours' flag blocks (peer-mapped)
    int wait_adv_event;    // > 0: before returning, also wait until every rank's adv_done >= this epoch
    int world;
};

__device__ __forceinline__ void copy_rows4(float4* dst, const float4* src, size_t n4, bool peer) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 4 * stride) {
        float4 t[4];  // loads back to back (NVLink latency for peer rows), then the stores
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i + k * stride < n4) t[k] = peer ? __ldcv(src + i + k * stride) : src[i + k * stride];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i + k * stride < n4) dst[i + k * stride] = t[k];
    }
}

// One halo exchange event, single handshake.  grid.x = CTAs per job, grid.y = njobs.
//  A. every CTA copies its share of my boundary rows into my outbox; the last one to finish
//     announces (release, system scope) that outbox[event & 1] is ready;
//  B. every CTA waits for the neighbour's announcement and pulls its share of the neighbour's
//     outbox rows into my halo rows.
// No "done reading" handshake is needed: outbox[event & 1] is rewritten at event + 2, which I can
// only reach after the neighbour announced event + 1, i.e. after its pull of this event completed.
__global__ void k_slab_pull(const __grid_constant__ PullArgs A) {
    __shared__ int s_ok;
    SlabFlags* mine = A.mine;
    const PullJob j = A.job[blockIdx.y];
    const int d = blockIdx.y & 1;
    if (j.out_rows > 0)
        copy_rows4(reinterpret_cast<float4*>(j.outbox), reinterpret_cast<const float4*>(j.mine), (size_t)j.out_rows * A.pitch / 4, false);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int total = gridDim.x * gridDim.y;
        if (atomicAdd(&mine->pull_done, 1) == total - 1) {
            mine->pull_done = 0;
            for (int dd = 0; dd < 2; ++dd)
                if (A.nbr[dd] >= 0) st_release_sys(&A.peer[dd]->ready[A.me], A.event);
        }
        s_ok = (A.nbr[d] < 0) ? 0 : (spin_until(&mine->ready[A.nbr[d]], A.event, &mine->error) ? 1 : 0);
    }
    __syncthreads();
    if (s_ok && j.in_rows > 0)
        copy_rows4(reinterpret_cast<float4*>(j.dst), reinterpret_cast<const float4*>(j.src), (size_t)j.in_rows * A.pitch / 4, true);
    // Folded k_slab_wait_all: whatever follows this exchange may overwrite fields that other ranks gathered
    // from during their advect; the flags were broadcast long ago, so this normally costs one load each.
    if (A.wait_adv_event > 0 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < A.world && (int)threadIdx.x != A.me)
        spin_until(&mine->adv_done[threadIdx.x], A.wait_adv_event, &mine->error);
}

// "Lazy" all-rank synchronisation around the peer gathers of advect: instead of two barriers,
// every rank BROADCASTS an epoch (no waiting) when it has finished writing the gathered fields
// (src_done, by k_advect4 pass 1) and when it has finished its advect (adv_done, by the last block of
// pass 2); a gathering thread waits only for the OWNER of a remote tap (pass 2), and a rank waits for
// everybody's adv_done only right before it overwrites the gathered fields -- inside the next halo
// exchange (k_slab_pull), or with k_slab_wait_all ahead of an upload / fill / inject.
struct SignalArgs {
    int event, me, world, which;   // which: 0 = src_done, 1 = adv_done
    SlabFlags* mine;
    SlabFlags* peer[SLAB_MAX_RANKS];
};
__global__ void k_slab_wait_all(const __grid_constant__ SignalArgs A) {
    const int t = threadIdx.x;
    if (t < A.world && t != A.me)
        spin_until(A.which == 0 ? &A.mine->src_done[t] : &A.mine->adv_done[t], A.event, &A.mine->error);
}

struct BarrierArgs {
    int event, me, world;
    SlabFlags* mine;
    SlabFlags* peer[SLAB_MAX_RANKS];
};
// All-ranks barrier (before / after the peer gathers of advect).
__global__ void k_slab_barrier(const __grid_constant__ BarrierArgs A) {
    const int t = threadIdx.x;
    if (t < A.world && t != A.me) {
        st_release_sys(&A.peer[t]->bar[A.me], A.event);
        spin_until(&A.mine->bar[t], A.event, &A.mine->error);
    }
}

}  // namespace fruid
