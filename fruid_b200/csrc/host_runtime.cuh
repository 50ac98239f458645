// host_runtime.cuh -- host-side plumbing shared by everything above the kernels: the thread-local error
// message behind fruid_last_error(), the CU / TRY / LAUNCH macros, the kernel-launch counter and the optional
// per-kernel CUDA-event timing that bench.py's roofline leg reads through fruid_profile_report().
#pragma once
#include "../../include/fruid_b200.h"
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>
#include "common.cuh"

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
// cudaStreamWaitEvent + the bookkeeping of programmatic launch chains (common.cuh): the next kernel on `st` has a
// dependency that is not its predecessor kernel, so it is launched without the programmatic attribute.
static inline cudaError_t stream_wait(cudaStream_t st, cudaEvent_t ev) {
    pdl_chain_break(st);
    return cudaStreamWaitEvent(st, ev, 0);
}
static inline int launched(const char* what, cudaStream_t st) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    (void)pdl_chain_continue(st);  // a kernel is now the newest operation on st (plain <<<>>> launches pass here too)
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
