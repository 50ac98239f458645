// field_set.cuh -- what a handle owns: the single device allocation [flags][fields][outbox], its streams and events,
// the slab (multi-GPU) bookkeeping with CUDA IPC mapping and halo exchange, staged asynchronous transfers, and the
// captured step graphs.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <new>
#include "operators.cuh"

using namespace fruid;

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
    int ev_adv_waited = 0;               // newest advect epoch whose adv_done of EVERY rank this stream has waited for
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
    std::vector<uint64_t> extra;   // unique ids (FieldSet::uid) of every handle whose buffers the graph bakes in
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
// Alternative halo exchange on the copy engines + stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32,
// resolved through the runtime's driver entry points like the tensor-map encoder): FRUID_B200_SLAB_EXCHANGE=ce.  Same
// protocol and the same bits as the exchange kernel, but measured SLOWER on 2 GPUs (1.251 vs 1.196 ms per 4096 x 8192
// step: sixteen small in-stream copies per exchange cost more than the SM slots the kernel has to wait for), so the
// default stays k_slab_pull.
struct StreamMemOps {
    typedef CUresult (*PFN_write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    typedef CUresult (*PFN_wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    PFN_write32 write32 = nullptr;
    PFN_wait32 wait32 = nullptr;
    bool ok = false;
};
static const StreamMemOps& stream_mem_ops() {
    static StreamMemOps M;
    static std::once_flag once;
    std::call_once(once, [] {
        void *pw = nullptr, *pq = nullptr;
        cudaDriverEntryPointQueryResult q1, q2;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &pw, cudaEnableDefault, &q1) == cudaSuccess && q1 == cudaDriverEntryPointSuccess &&
            cudaGetDriverEntryPoint("cuStreamWaitValue32", &pq, cudaEnableDefault, &q2) == cudaSuccess && q2 == cudaDriverEntryPointSuccess) {
            M.write32 = reinterpret_cast<StreamMemOps::PFN_write32>(pw);
            M.wait32 = reinterpret_cast<StreamMemOps::PFN_wait32>(pq);
            M.ok = M.write32 && M.wait32;
        } else {
            cudaGetLastError();
        }
    });
    return M;
}
static std::atomic<uint64_t> g_exchanges_on_copy_engines{0};
static int g_slab_exchange_mode = -1;  // -1 = not decided yet, 0 = kernel, 1 = copy engines
static bool slab_exchange_on_copy_engines() {
    if (g_slab_exchange_mode < 0) {
        const char* e = getenv("FRUID_B200_SLAB_EXCHANGE");
        g_slab_exchange_mode = (e && strcmp(e, "ce") == 0 && stream_mem_ops().ok) ? 1 : 0;
    }
    return g_slab_exchange_mode == 1;
}

// Process-unique handle ids: a graph remembers the ids (not the host addresses) of the handles it was captured
// with, so a handle that malloc places at the address of a destroyed one can never replay the dead handle's graph.
static std::atomic<uint64_t> g_next_uid{1};

struct FieldSet {
    uint64_t uid = g_next_uid.fetch_add(1);
    StepGraph graph;
    StepGraph batch_graph;                       // fruid_density_step_dev_batch led by this handle
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;  // fork / join of the batched step (also inside captures)
    Geom g{};
    int n = 0;       // buffers allocated: user fields, then scratch (lib.rs:213,252), then the add_source temporary
    int n_user = 0;  // fields the crate API exposes
    float* buf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    char* base = nullptr;  // the one device allocation: [flag block][field 0][field 1]...
    size_t field_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // velocity handles: the v diffuse runs here beside the u diffuse (buf[6], buf[7])
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
    // Output buffer of the draw_* entry points, kept between calls: a stream-ordered allocation per frame would be
    // returned to the OS at every synchronisation (default pool threshold) and cost ~0.4 ms per call to get back.
    float* draw_buf = nullptr;
    size_t draw_bytes = 0;
    float* vel_stage[2] = {nullptr, nullptr};  // fruid_density_step_host: device copies of host-side (u, v)
    int draw_buffer(size_t bytes, float** out) {
        if (bytes > draw_bytes) {
            if (draw_buf) { CU(cudaStreamSynchronize(stream)); CU(cudaFree(draw_buf)); draw_buf = nullptr; draw_bytes = 0; }
            CU(cudaMalloc(&draw_buf, bytes));
            draw_bytes = bytes;
        }
        *out = draw_buf;
        return 0;
    }
    // Host mirrors (fruid_fluid_bind_mirror): the caller's arrays that shadow u and v.  A mirror is current while
    // mirror_epoch == dev_epoch, i.e. it was downloaded after the last change of the device fields; then
    // fruid_density_step_host(d, u, v, ..) with exactly these pointers can use the device-resident velocity.
    uint64_t dev_epoch = 1;
    const float* mirror[4] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t mirror_epoch[4] = {0, 0, 0, 0};
    void touch() { ++dev_epoch; }
    int iters = 20;  // src/lib.rs:52
    int fused = 0;   // 0 = library default
    int device = 0;
    SlabInfo slab;
    // Slab mode: events of other handles' streams (density steps) that still READ my u, v; awaited
    // lazily right before the step first overwrites u or v instead of at the step's start.
    std::vector<cudaEvent_t> readers;
    int wait_readers() {
        TRY(quiesce_gathers());
        for (cudaEvent_t e : readers) CU(stream_wait(stream, e));
        readers.clear();
        return 0;
    }
    // Slab mode: other ranks gather from my fields during their advect (peer loads).  Before anything from
    // outside the step schedule writes a field (upload, fill, inject), wait until every rank has finished the
    // last advect; inside the schedule the same wait rides in a halo exchange (exchange(.., wait_adv_event)).
    int quiesce_gathers() {
        if (!is_slab() || slab.ev_adv <= slab.ev_adv_waited) return 0;
        TRY(wait_all(1, slab.ev_adv));
        slab.ev_adv_waited = slab.ev_adv;
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

    int init(size_t w, size_t gh, int nfields, int rank = 0, int world = 1, int extra = 0) {
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
        n = nfields + extra;
        n_user = nfields - 2;
        CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        if (extra) CU(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
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
        if (draw_buf) cudaFree(draw_buf);
        for (float* p : vel_stage) if (p) cudaFree(p);
        if (dl_snap) cudaEventDestroy(dl_snap);
        if (dl_done) cudaEventDestroy(dl_done);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (base) cudaFree(base);
        if (slab.adv_block_flag) cudaFree(slab.adv_block_flag);
        if (ev) cudaEventDestroy(ev);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (stream2) { cudaStreamSynchronize(stream2); tile_sched_release(stream2); cudaStreamDestroy(stream2); }
        if (stream) { tile_sched_release(stream); cudaStreamDestroy(stream); }
    }
    // Host arrays cover the rows this rank OWNS (the whole field on one GPU), dense w floats per row.
    int upload(int field, const float* host) {
        if (field < 0 || field >= n_user || !host) return fail(FRUID_ERR_BAD_ARG, "bad field %d or null host pointer", field);
        TRY(wait_readers());
        touch();
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
        if (host == mirror[field]) mirror_epoch[field] = dev_epoch;
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
        touch();
        const size_t rows = (size_t)(own_hi() - own_lo());
        if (!up_stage[field]) {
            CU(cudaMalloc(&up_stage[field], g.pitch * rows * sizeof(float)));
            CU(cudaEventCreateWithFlags(&up_ev[field], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&up_used[field], cudaEventDisableTiming));
        } else {
            CU(stream_wait(copy_stream, up_used[field]));  // previous staged upload has been consumed
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
            CU(stream_wait(stream, up_ev[f]));
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
            CU(stream_wait(stream, dl_done));  // previous download has left the snapshot buffer
        }
        CU(cudaMemcpyAsync(dl_stage, buf[field] + (size_t)own_lo() * g.pitch, g.pitch * rows * sizeof(float),
                           cudaMemcpyDeviceToDevice, stream));
        CU(cudaEventRecord(dl_snap, stream));
        CU(stream_wait(copy_stream, dl_snap));
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
        touch();
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
        if (wait_adv_event > slab.ev_adv_waited) slab.ev_adv_waited = wait_adv_event;
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
        if (slab_exchange_on_copy_engines()) {
            // The same single-handshake protocol as k_slab_pull (outboxes alternating by event parity), carried by the
            // copy engines and stream memory operations instead of a kernel: no SM has to be free for it (the
            // persistent tile kernel of the other stream owns every SM's register file for ~85 us at a time).
            const StreamMemOps& M = stream_mem_ops();
            for (int j = 0; j < nj; ++j) {
                const PullJob& J = A.job[j];
                if (J.out_rows > 0)
                    CU(cudaMemcpyAsync(J.outbox, J.mine, (size_t)J.out_rows * g.pitch * sizeof(float), cudaMemcpyDeviceToDevice, stream));
            }
            for (int d = 0; d < 2; ++d)   // announce: my outbox[par] holds event A.event (prior copies are flushed first)
                if (A.nbr[d] >= 0 && M.write32(stream, (CUdeviceptr)&A.peer[d]->ready[me], (cuuint32_t)A.event, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS)
                    return fail(FRUID_ERR_CUDA, "cuStreamWriteValue32 failed (halo exchange)");
            for (int d = 0; d < 2; ++d) {
                if (A.nbr[d] < 0) continue;
                if (M.wait32(stream, (CUdeviceptr)&A.mine->ready[A.nbr[d]], (cuuint32_t)A.event, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                    return fail(FRUID_ERR_CUDA, "cuStreamWaitValue32 failed (halo exchange)");
                for (int j = d; j < nj; j += 2) {
                    const PullJob& J = A.job[j];
                    if (J.in_rows > 0)
                        CU(cudaMemcpyAsync(J.dst, J.src, (size_t)J.in_rows * g.pitch * sizeof(float), cudaMemcpyDeviceToDevice, stream));
                }
            }
            if (wait_adv_event > 0)
                for (int r = 0; r < slab.world; ++r)
                    if (r != me && M.wait32(stream, (CUdeviceptr)&A.mine->adv_done[r], (cuuint32_t)wait_adv_event, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                        return fail(FRUID_ERR_CUDA, "cuStreamWaitValue32 failed (advect epoch)");
            pdl_chain_break(stream);
            g_exchanges_on_copy_engines.fetch_add(1, std::memory_order_relaxed);
            return 0;
        }
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

struct fruid_fluid { FieldSet s; };    // buf: 0=u 1=v 2=u_prev 3=v_prev 4=scratch (lib.rs:247-253) 5=add_source temporary; single GPU: 6, 7 = the same two for the concurrent v diffuse
struct fruid_density { FieldSet s; };  // buf: 0=dens 1=dens_prev 2=scratch (lib.rs:210-214) 3=add_source temporary


// Runs `body` (which enqueues one step on S.stream) either directly or as a replayed CUDA graph.
// A graph is captured on the second step with given parameters (the first warms lazy allocations),
// and only kept if the step left the buffer roles unchanged (an even number of launches per
// lin_solve), because the graph bakes the pointers in.
template <class Body>
static int run_graphed(StepGraph& G, FieldSet& S, const std::vector<FieldSet*>& parts, float dt, float p, const void* a0,
                       const void* a1, const std::vector<uint64_t>& extra, Body body) {
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
        pdl_chain_break(S.stream);
        for (FieldSet* P : parts) { pdl_chain_break(P->stream); if (P->stream2) pdl_chain_break(P->stream2); }
        g_launches.fetch_add(G.kernels, std::memory_order_relaxed);
        return 0;
    }
    if (G.failed) return body();
    // second step with the same parameters: capture it
    std::vector<float*> before;
    for (FieldSet* P : parts) before.insert(before.end(), P->buf, P->buf + 8);
    auto restore = [&]() { size_t k = 0; for (FieldSet* P : parts) for (int i = 0; i < 8; ++i) P->buf[i] = before[k++]; };
    auto unchanged = [&]() { size_t k = 0; bool same = true; for (FieldSet* P : parts) for (int i = 0; i < 8; ++i) same &= (P->buf[i] == before[k++]); return same; };
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
    pdl_chain_break(S.stream);
    for (FieldSet* P : parts) { pdl_chain_break(P->stream); if (P->stream2) pdl_chain_break(P->stream2); }
    g_launches.fetch_add(nk, std::memory_order_relaxed);
    return 0;
}
template <class Body>
static int run_step_graphed(FieldSet& S, float dt, float p, const void* a0, const void* a1, Body body, uint64_t vel_uid = 0) {
    return run_graphed(S.graph, S, {&S}, dt, p, a0, a1, {S.uid, vel_uid}, body);
}
