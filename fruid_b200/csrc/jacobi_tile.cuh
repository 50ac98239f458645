// jacobi_tile.cuh -- host side of the fused Jacobi kernel (jacobi_tile_kernel.cuh): tile shape and fused depth
// per launch, tensor maps, per-stream scheduler words, the verified reciprocal division, launch dispatch.
#pragma once
#include <cuda.h>

#include <cstring>
#include <iterator>
#include <mutex>
#include <unordered_map>

#include "jacobi_tile_kernel.cuh"

namespace fruid {

// Exhaustive check of the MODE_RECIP division for one divisor: all 2^23 significands, two binades, both signs.
__global__ void k_recip_verify(float c, float zh, float zl, int variant, int* bad) {
    const unsigned m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (1u << 23)) return;
    int b = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned bits = ((k & 1) ? 0x42000000u : 0x3f800000u) | ((k & 2) ? 0x80000000u : 0u) | m;
        const float v = __uint_as_float(bits);
        float q;
        if (variant == MODE_RECIP) {
            q = __fmaf_rn(v, zh, __fmul_rn(v, zl));
        } else {  // MODE_RECIP3
            const float q0 = __fmul_rn(v, zh);
            q = __fmaf_rn(__fmaf_rn(q0, -c, v), zh, q0);
        }
        b |= (__float_as_uint(q) != __float_as_uint(__fdiv_rn(v, c)));
    }
    if (b) atomicOr(bad, 1);
}

// ---- host side -----------------------------------------------------------------------------
// c -> (usable, zh, zl); filled by recip_for() the first time a divisor is seen (synchronises once).
struct RecipInfo { bool ok; int mode; float zh, zl; };
static std::mutex g_recip_mu;
static std::unordered_map<uint32_t, RecipInfo> g_recip;
static bool recip_for(float c, RecipInfo* out, cudaStream_t st) {
    uint32_t key;
    memcpy(&key, &c, 4);
    std::lock_guard<std::mutex> lk(g_recip_mu);
    auto it = g_recip.find(key);
    if (it != g_recip.end()) { *out = it->second; return it->second.ok; }
    RecipInfo r{false, MODE_GENERAL, 0.f, 0.f};
    int e = 0;
    (void)frexpf(c, &e);
    if (c == c && c != 0.f && e > -30 && e < 30) {  // finite, normal, moderate magnitude
        // Candidates, cheapest first: the two-instruction form q = fma(v, zh, v*zl) with the double-word reciprocal
        // (zh, zl) -- correctly rounded for "almost all" c; when one significand in 2^23 rounds the wrong way, a
        // neighbouring double-word (zl one ulp up or down, or zh one ulp off with zl re-derived) often has no failing
        // significand at all -- then the three-instruction form with the nearest reciprocal.  Whatever is used has
        // passed the exhaustive comparison with __fdiv_rn.
        const float zh0 = 1.0f / c;
        const double inv = 1.0 / (double)c;
        struct Cand { int variant; float zh, zl; };
        Cand cands[6];
        int nc = 0;
        const float zl0 = (float)(inv - (double)zh0);
        cands[nc++] = {(int)MODE_RECIP, zh0, zl0};
        cands[nc++] = {(int)MODE_RECIP, zh0, nextafterf(zl0, INFINITY)};
        cands[nc++] = {(int)MODE_RECIP, zh0, nextafterf(zl0, -INFINITY)};
        for (float dir : {INFINITY, -INFINITY}) {
            const float zh1 = nextafterf(zh0, dir);
            cands[nc++] = {(int)MODE_RECIP, zh1, (float)(inv - (double)zh1)};
        }
        cands[nc++] = {(int)MODE_RECIP3, zh0, zl0};
        // The check runs on the LAUNCHING stream (the caller made sure it is not capturing) and is followed by a
        // synchronisation of that stream only.
        int* bad = nullptr;
        if (cudaMalloc(&bad, sizeof(int)) == cudaSuccess) {
            for (int k = 0; k < nc; ++k) {
                int h = 1;
                if (cudaMemsetAsync(bad, 0, sizeof(int), st) != cudaSuccess) break;
                k_recip_verify<<<(1u << 23) / 256, 256, 0, st>>>(c, cands[k].zh, cands[k].zl, cands[k].variant, bad);
                if (cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
                if (cudaStreamSynchronize(st) != cudaSuccess) break;
                if (h == 0) {
                    r.ok = true; r.mode = cands[k].variant; r.zh = cands[k].zh; r.zl = cands[k].zl;
                    break;
                }
            }
            cudaFree(bad);
        }
        pdl_chain_break(st);
    }
    g_recip.emplace(key, r);
    *out = r;
    return r.ok;
}

// Process-wide tuning knobs (fruid_set_tile_config).  Read as one snapshot per launch decision, written under
// the same mutex: handles may be used from different threads.
struct TileCfg { int R, NW; };
// Default shape 7 x 20 (128 x 140 cells, 640 threads): the fastest of the shapes that fit the register file at
// 4096^2 (8.24 us per sweep vs 8.35-8.45 for 8 x 16; profiles/tune_tile_r2.json) -- 17 % redundant rows instead of 19 %.
struct TileKnobs { TileCfg cfg = {7, 20}; int default_T = 10; bool automatic = true; };
static std::mutex g_tile_mu;
static TileKnobs g_tile_knobs;
static inline TileKnobs tile_knobs() { std::lock_guard<std::mutex> lk(g_tile_mu); return g_tile_knobs; }

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

struct TmKey {
    const void* p; size_t pitch; int h; int th; int tw;
    bool operator==(const TmKey& o) const { return p == o.p && pitch == o.pitch && h == o.h && th == o.th && tw == o.tw; }
};
struct TmKeyHash {
    size_t operator()(const TmKey& k) const {
        return std::hash<const void*>()(k.p) ^ (k.pitch * 1315423911u) ^ ((size_t)k.h << 20) ^ (size_t)k.th ^ ((size_t)k.tw << 10);
    }
};
static std::mutex g_tm_mu;
static std::unordered_map<TmKey, CUtensorMap, TmKeyHash> g_tm_cache;

// Tensor map of a field for (tw x th) boxes, zero fill outside.  Returns 0 or -4 (CUDA error).
static int field_tensor_map(const float* base, size_t pitch, int h, int tw, int th, CUtensorMap* out) {
    std::lock_guard<std::mutex> lk(g_tm_mu);
    const TmKey key{base, pitch, h, th, tw};
    auto it = g_tm_cache.find(key);
    if (it != g_tm_cache.end()) { *out = it->second; return 0; }
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return -4;
    const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)h};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)tw, (cuuint32_t)th};
    const cuuint32_t estr[2] = {1u, 1u};
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -4;
    if (g_tm_cache.size() > 4096) g_tm_cache.clear();
    g_tm_cache.emplace(key, tm);
    *out = tm;
    return 0;
}
static inline int tile_tensor_map(const float* base, size_t pitch, int h, int th, CUtensorMap* out) {
    return field_tensor_map(base, pitch, h, 128, th, out);  // the (128 x th) boxes of the fused Jacobi kernel
}

// Per-stream scheduler words (zeroed once; every launch leaves them zero again).
static std::mutex g_sched_mu;
static std::unordered_map<cudaStream_t, int*> g_sched;
static int* tile_sched_for(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_sched_mu);
    auto it = g_sched.find(st);
    if (it != g_sched.end()) return it->second;
    int* p = nullptr;
    if (cudaMalloc(&p, 2 * sizeof(int)) != cudaSuccess) return nullptr;
    // Zeroed ON THE STREAM that will use them: cudaMemset would run on the legacy default stream, which does not
    // order with non-blocking streams -- the first kernel could read the words before they are cleared.
    if (cudaMemsetAsync(p, 0, 2 * sizeof(int), st) != cudaSuccess) { cudaFree(p); return nullptr; }
    pdl_chain_break(st);
    g_sched.emplace(st, p);
    return p;
}

static inline int tiles_needed(int extent, int tile, int stride) {
    if (extent <= tile) return 1;
    return (extent - tile + stride - 1) / stride + 1;
}

static void tile_sched_release(cudaStream_t st) {  // the stream is going away (handle destroyed)
    std::lock_guard<std::mutex> lk(g_sched_mu);
    auto it = g_sched.find(st);
    if (it == g_sched.end()) return;
    cudaFree(it->second);
    g_sched.erase(it);
}

static inline int jacobi_tile_max_T(int R, int NW) {
    const int th = R * NW;
    int H = (th < 128 ? th : 128) / 2 - 2;  // keep at least 4 output rows/cols per tile
    return H & ~1;
}

static inline long jacobi_tile_count(const Geom& g, int R, int NW, int T) {
    const int H = T + (T & 1), TH = R * NW;
    return (long)tiles_needed(g.w, 128, 128 - 2 * H) * tiles_needed(g.h, TH, TH - 2 * H);
}

// Tile shape and fused depth per launch.  The kernel's cost per tile and sweep is proportional to the
// tile's rows (it is issue bound), so as long as a tiling leaves SMs idle a SMALLER tile finishes sooner.
// Measured scene step (tools/small_grid_cfg_probe.py, us; 128x128 tiles at T = 20 in brackets):
//   250^2: 86 with 128x32 tiles (162)    512^2: 104 with 128x48 (167)
//   768^2: 119 with 128x64 (169)         1024^2: 156 with 128x96 (175)       all at T = 10.
// From ~1200^2 upwards every SM has a tile anyway and the least halo recompute wins: 128x128, T = 10.
// Automatic choice (until fruid_set_tile_config pins a shape; always the pinned / default shape in slab mode,
// whose halo bookkeeping assumes it): the first of these that gives every tile its own SM.
struct TilePlan { int R, NW, T; };
static const TilePlan kSmallGridPlans[] = {{2, 16, 10}, {3, 16, 10}, {2, 32, 10}, {4, 24, 10}, {8, 16, 20}};  // then the default shape

static inline int jacobi_tile_default_fused(const Geom& g, int iters) {
    const TileKnobs K = tile_knobs();
    const TileCfg g_tile_cfg = K.cfg;
    int T = K.default_T;
    const bool whole = (g.gh == g.h);
    if (whole && K.automatic) {
        for (const TilePlan& p : kSmallGridPlans) {
            const int t = p.T > iters ? iters : p.T;
            if (jacobi_tile_count(g, p.R, p.NW, t) <= 148) return t < 1 ? 1 : t;
        }
    } else if (whole) {
        // pinned shape: when the tiling at the default depth cannot even give every SM one tile, fuse a
        // whole 20-sweep solve per launch (more halo recompute on SMs that would otherwise idle, half the launches)
        if (jacobi_tile_count(g, g_tile_cfg.R, g_tile_cfg.NW, T) < 148) T = 20;
    }
    const int mx = jacobi_tile_max_T(g_tile_cfg.R, g_tile_cfg.NW);
    if (T > mx) T = mx;
    if (T > iters) T = iters;
    return T < 1 ? 1 : T;
}

// Grids of a few tiles per SM: the tile count quantises badly over 148 SMs (195 tiles of 128 x 140 at 1536^2 are two
// rounds, the second a third full), and which shape wastes least depends on the size.  Measured 20-sweep Poisson solves
// (tools/tune_tile.py --size; ms for 7x20 / this shape): 1536^2 0.054 / 0.048 (6x20), 2048^2 0.062 / 0.058 (8x16),
// 3072^2 0.108 / 0.103 (10x16); from ~3500^2 on the default shape wins (4096^2: 8.7 rounds).
static inline TileCfg jacobi_tile_midsize_shape(const Geom& g, const TileCfg& dflt) {
    const size_t cells = (size_t)g.w * (size_t)g.h;
    if (cells < (size_t)1800 * 1800) return TileCfg{6, 20};
    if (cells < (size_t)2600 * 2600) return TileCfg{8, 16};
    if (cells < (size_t)3500 * 3500) return TileCfg{10, 16};
    return dflt;
}

// Shape for a launch of T fused sweeps.
static inline TileCfg jacobi_tile_shape(const Geom& g, int T) {
    const TileKnobs K = tile_knobs();
    const TileCfg g_tile_cfg = K.cfg;
    if (K.automatic && g.gh == g.h) {
        for (const TilePlan& p : kSmallGridPlans)
            if (T <= jacobi_tile_max_T(p.R, p.NW) && jacobi_tile_count(g, p.R, p.NW, T) <= 148) return TileCfg{p.R, p.NW};
        const TileCfg m = jacobi_tile_midsize_shape(g, g_tile_cfg);
        if (T <= jacobi_tile_max_T(m.R, m.NW)) return m;
    }
    return g_tile_cfg;
}

// SM count per device (handles may live on different devices of one process).
static int device_num_sms() {
    static std::mutex mu;
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    std::lock_guard<std::mutex> lk(mu);
    if (sms[dev] == 0 && cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms[dev] = 0;
    return sms[dev];
}

#define FRUID_TILE_CONFIGS(X) X(8, 16) X(10, 16) X(6, 16) X(6, 20) X(7, 20) X(5, 24) X(4, 24) X(8, 12) X(4, 16) X(2, 32) X(2, 16) X(4, 8) X(3, 16)

// Returns 0 or a negative fruid_status (-7 = bad argument, -4 = CUDA error).
static int jacobi_tile_launch(float* out, const float* x, const float* x0, const Geom& g, float a, float c, int b, int T,
                              int x_is_zero, cudaStream_t st, int fuse_src = 0, float src_dt = 0.f, float* z_out = nullptr) {
    if (T < 1) return -7;
    const TileCfg cfg = jacobi_tile_shape(g, T);
    const int R = cfg.R, NW = cfg.NW, TH = R * NW;
    if (T > jacobi_tile_max_T(R, NW)) return -7;
    TileArgs A;
    if (int r = tile_tensor_map(x0, g.pitch, g.h, TH, &A.tm_x0)) return r;
    if (int r = tile_tensor_map(x, g.pitch, g.h, TH, &A.tm_x)) return r;
    A.out = out; A.x = x; A.g = g; A.a = a; A.b = b; A.T = T; A.x_is_zero = x_is_zero;
    A.zh = A.zl = 0.f;
    A.fuse_src = fuse_src; A.src_dt = src_dt; A.z_out = z_out;
    A.sched = tile_sched_for(st);
    if (!A.sched) return -4;
    A.H = T + (T & 1);
    A.SX = 128 - 2 * A.H;
    A.SY = TH - 2 * A.H;
    A.ntx = tiles_needed(g.w, 128, A.SX);
    A.nty = tiles_needed(g.h, TH, A.SY);
    int mode = MODE_GENERAL;
    A.c = c;
    if (c == 1.0f) {
        mode = (a == 0.0f) ? MODE_ZEROA : MODE_ONE;
    } else {
        int e = 0;
        const float m = frexpf(c, &e);
        if (m == 0.5f && e > -100 && e < 100) {  // c = 2^k: multiply by the exact reciprocal
            mode = (a == 1.0f) ? MODE_POISSON : MODE_POW2;
            A.c = 1.0f / c;
        } else {
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            RecipInfo ri;
            const bool capturing = cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone;
            bool known;
            {
                uint32_t key; memcpy(&key, &c, 4);
                std::lock_guard<std::mutex> lk(g_recip_mu);
                known = g_recip.count(key) != 0;
            }
            // never synchronise inside a stream capture: an unseen divisor takes the IEEE division there
            if ((known || !capturing) && recip_for(c, &ri, st)) { mode = ri.mode; A.zh = ri.zh; A.zl = ri.zl; }
        }
    }
    const int num_sms = device_num_sms();
    if (num_sms <= 0) return -4;
    A.pdl_trigger = g_pdl.enabled ? 1 : 0;
    A.dbg_delay_ns = g_dbg_tail_delay_host;
    const bool pdl = g_pdl.enabled && (pdl_chain_continue(st) || g_pdl.unsafe);
    int r = jacobi_tile_launch_inst_a(R, NW, mode, A, num_sms, pdl, st);
    if (r == -1) r = jacobi_tile_launch_inst_b(R, NW, mode, A, num_sms, pdl, st);
    if (r == -1) r = jacobi_tile_launch_inst_c(R, NW, mode, A, num_sms, pdl, st);
    if (r == -1) r = jacobi_tile_launch_inst_d(R, NW, mode, A, num_sms, pdl, st);
    if (r == -1) return -7;
    return r == 0 ? 0 : -4;
}

static inline int jacobi_tile_set_cfg(int R, int NW, int T) {
    std::lock_guard<std::mutex> lk(g_tile_mu);
    if (R == 0 && NW == 0) {  // back to the automatic choice
        g_tile_knobs.automatic = true; g_tile_knobs.cfg = {7, 20}; g_tile_knobs.default_T = T > 0 ? T : 10;
        return 0;
    }
#define FRUID_TILE_OK(r, nw) if (R == r && NW == nw) { g_tile_knobs.cfg = {R, NW}; g_tile_knobs.automatic = false; if (T > 0) g_tile_knobs.default_T = T; return 0; }
    FRUID_TILE_CONFIGS(FRUID_TILE_OK)
#undef FRUID_TILE_OK
    return -7;
}

}  // namespace fruid
