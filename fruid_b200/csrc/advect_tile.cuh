// advect_tile.cuh -- advect (src/lib.rs:84-126) with the gathered fields staged in shared memory.
//
// k_advect4 (kernels_vec.cuh) gathers its 4 bilinear taps per cell and field with scalar global loads; with one
// float4 of cells per lane the 32 lanes of a tap load are 16 bytes apart, four L1 wavefronts per instruction, and the
// kernel is bound by exactly that (ncu: LSU issue, 0.48 of the HBM roofline at 4096^2).  Here a block of 256 threads
// produces 128 columns x 32 rows:
//   * one thread issues a TMA load (cp.async.bulk.tensor.2d, zero fill outside the field) of the 144 x 40 window of
//     every gathered field around the block -- 8 columns / 4 rows of margin -- so HBM and L2 see one coalesced read of
//     each field (window / block = 1.41x, the overlap is served by L2);
//   * lane l of a warp owns columns l, l + 32, l + 64, l + 96 of four rows: the taps of neighbouring lanes are
//     neighbouring words, one shared-memory wavefront per tap instruction, with the row stride of the window as an
//     immediate offset (one address computation per cell for all taps of all fields);
//   * a thread whose four cells may trace back beyond the window (a displacement above 7 columns or 3 rows) gathers
//     from global memory instead -- same taps, same arithmetic, so results never depend on which path ran;
//   * velocity self-advection (lib.rs:204-205: the gathered fields ARE the advecting velocity) reads the cell's own
//     velocity from the window too, i.e. u0, v0 are read once; a density advect gets u, v as two more TMA boxes;
//   * set_bnd (lib.rs:125) is folded into the stores; only blocks that touch the domain border pay for its tests.
// Arithmetic is the scalar code of kernels_basic.cuh (backtrace(), mixf()): bit-identical to k_advect / k_advect4.
#pragma once
#include "jacobi_tile.cuh"
#include "kernels_basic.cuh"
#include "kernels_vec.cuh"

namespace fruid {

constexpr int ADV_TW = 128, ADV_TH = 32;          // cells produced per block
constexpr int ADV_MX = 8, ADV_MY = 4;             // window margins (columns: a multiple of 4 keeps rows 16-byte aligned)
constexpr int ADV_WW = ADV_TW + 2 * ADV_MX;       // 144 floats per window row (576 B)
constexpr int ADV_WH = ADV_TH + 2 * ADV_MY;       // 40 rows
constexpr int ADV_FIELD_FLOATS = ADV_WW * ADV_WH; // 5760 floats = 23040 B per field

struct AdvectTileArgs {
    CUtensorMap tm[ADVECT_MAX_FIELDS];  // gathered fields d0[f]: dims (pitch, h), box (ADV_WW, ADV_WH)
    CUtensorMap tm_u, tm_v;             // the advecting velocity, box (ADV_TW, ADV_TH) (UVT kernels only)
    AdvectArgs a;
    const float* u;
    const float* v;
    Geom g;
    float dt0, dt1, xmax, ymax;
    int ly0, ly1;                       // LOCAL rows to produce: [1, h-1) on one GPU, the owned rows of a slab
};

// store_with_bnd (common.cuh) for a y-slab: the cell sits at LOCAL row ly, boundary logic uses its global row.
__device__ __forceinline__ void store_with_bnd_rows(float* __restrict__ out, const Geom& g, int b, int i, int ly, float v) {
    const size_t k = at(g, i, ly);
    out[k] = v;
    const int gy = ly + g.gy0;
    const bool nxg = (b == B_NEGX), nyg = (b == B_NEGY);
    const bool xl = (i == 1), xr = (i == g.nx), yt = (gy == 1), yb = (gy == g.ny);
    if (xl | xr | yt | yb) {
        const float vx = sgn(nxg, v), vy = sgn(nyg, v);
        const size_t up = k - g.pitch, dn = k + g.pitch;
        if (xl) out[k - 1] = vx;
        if (xr) out[k + 1] = vx;
        if (yt) out[up] = vy;
        if (yb) out[dn] = vy;
        const float c = fmul(0.5f, fadd(vy, vx));
        if (xl && yt) out[up - 1] = c;
        if (xl && yb) out[dn - 1] = c;
        if (xr && yt) out[up + 1] = c;
        if (xr && yb) out[dn + 1] = c;
    }
}

// Where a cell's own velocity comes from: the windows themselves (velocity self-advection), two more TMA boxes
// (one or two gathered fields: everything arrives asynchronously, nothing waits on a register), or plain loads
// (three or four dyes: the windows already fill the shared memory an SM can give two blocks).
enum : int { ADV_UV_SELF = 0, ADV_UV_TMA = 1, ADV_UV_LDG = 2 };
constexpr int ADV_UV_FLOATS = ADV_TW * ADV_TH;    // 4096 floats = 16 KiB per velocity component
template <int NF, int UV>
constexpr size_t advect_tile_smem_bytes() {
    return ((size_t)NF * ADV_FIELD_FLOATS + (UV == ADV_UV_TMA ? 2 * ADV_UV_FLOATS : 0)) * sizeof(float) + 16;
}

// min / max that keep a NaN operand (FMNMX.NAN): `if (x < lo) x = lo; if (x > hi) x = hi;` of lib.rs:95-110 leaves
// a NaN alone (both comparisons are false), exactly like these.
__device__ __forceinline__ float max_keep_nan(float x, float lo) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(lo));
    return r;
}
__device__ __forceinline__ float min_keep_nan(float x, float hi) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(hi));
    return r;
}

// PASS 0: one GPU.  PASS 1: a y-slab (kernels_vec.cuh: k_advect4<NF, 1>) -- only quads of four aligned columns whose
// eight tap rows are all owned by this rank are produced here; the others queue their 128 x 8 block of k_advect4's
// geometry for its pass 2, which gathers them from the owners' memory.  A.a.d0[] is then indexed by GLOBAL row, the
// windows and everything else by local row.
template <int NF, int UV, int PASS>
__global__ void __launch_bounds__(256) k_advect_tile(const __grid_constant__ AdvectTileArgs A, const __grid_constant__ SlabTable tab) {
    pdl_enter();
    if (PASS == 1 && blockIdx.x == 0 && blockIdx.y == 0 && (int)threadIdx.x < tab.world)
        st_release_sys(&tab.flags[threadIdx.x]->src_done[tab.me], tab.event);  // folded "src_done" broadcast, see k_advect4
    extern __shared__ __align__(128) unsigned char adv_smem[];
    constexpr uint32_t TX_BYTES = (uint32_t)((NF * ADV_FIELD_FLOATS + (UV == ADV_UV_TMA ? 2 * ADV_UV_FLOATS : 0)) * sizeof(float));
    float* const win = reinterpret_cast<float*>(adv_smem);
    float* const uvt = win + NF * ADV_FIELD_FLOATS;  // ADV_UV_TMA: u tile, then v tile
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(adv_smem + TX_BYTES);
    const Geom& g = A.g;
    const int bx0 = blockIdx.x * ADV_TW;           // first column of the block (column 0 is the border)
    const int by0 = A.ly0 + blockIdx.y * ADV_TH;   // first row of the block (local)
    const int wx0 = bx0 - ADV_MX, wy0 = by0 - ADV_MY;  // window origin
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        mbar_expect_tx(mbar, TX_BYTES);
        if (UV == ADV_UV_TMA) {
            tma_load_2d(uvt, &A.tm_u, bx0, by0, mbar);
            tma_load_2d(uvt + ADV_UV_FLOATS, &A.tm_v, bx0, by0, mbar);
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) tma_load_2d(win + f * ADV_FIELD_FLOATS, &A.tm[f], wx0, wy0, mbar);
    }
    __syncthreads();  // the barrier is initialised before anybody waits on it
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    // set_bnd tests only where the block touches the border (block-uniform)
    const bool edge = (bx0 == 0) || (bx0 + ADV_TW >= g.nx) || (by0 + g.gy0 == 1) || (by0 + g.gy0 + ADV_TH > g.ny);
    const int i_first = bx0 + lane;                // my columns: i_first + 32 c
    float fi[4];
    bool col_ok[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        fi[c] = (float)(i_first + 32 * c);
        col_ok[c] = (i_first + 32 * c >= 1) && (i_first + 32 * c <= g.nx);
    }
    const bool deep = (PASS == 1) && (by0 + g.gy0 - ADV_MY >= tab.my_lo) && (by0 + g.gy0 + ADV_TH + ADV_MY < tab.my_hi);
    const int jw = by0 + wi * 4;                   // my first row
    const int nrows = min(4, A.ly1 - jw);          // warp-uniform
    // Window word of tap (i0, j0) = j0*ADV_WW + i0 - wbase (j0 a GLOBAL row); the centre cell (i, j) likewise.
    const int wbase = (wy0 + g.gy0) * ADV_WW + wx0;
    const float dt0 = A.dt0, dt1 = A.dt1, xmax = A.xmax, ymax = A.ymax;
    mbar_wait(mbar, 0);
    dbg_tail_delay(wi);
#pragma unroll 1
    for (int rr = 0; rr < nrows; ++rr) {
        const int jl = jw + rr, j = jl + g.gy0;    // local / global row
        const float fj = (float)j;
        const size_t krow = (size_t)jl * g.pitch + (size_t)i_first;
        float uc[4], vc[4];
        if (UV == ADV_UV_SELF) {
            const float* const wrow = win + (j * ADV_WW + i_first - wbase);
#pragma unroll
            for (int c = 0; c < 4; ++c) { uc[c] = wrow[32 * c]; vc[c] = wrow[32 * c + ADV_FIELD_FLOATS]; }
        } else if (UV == ADV_UV_TMA) {
            const float* const urow = uvt + (jl - by0) * ADV_TW + lane;
#pragma unroll
            for (int c = 0; c < 4; ++c) { uc[c] = urow[32 * c]; vc[c] = urow[32 * c + ADV_UV_FLOATS]; }
        } else {
            const float* const urow = A.u + krow;
            const float* const vrow = A.v + krow;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uc[c] = col_ok[c] ? __ldg(urow + 32 * c) : 0.f;
                vc[c] = col_ok[c] ? __ldg(vrow + 32 * c) : 0.f;
            }
        }
        // Displacements in cells (lib.rs:92-93).  |dt0*u| <= 7 and |dt1*v| <= 3 for all four cells puts every tap inside
        // the window: x = i - dt0*u rounds monotonically, so i0 = trunc(clamped x) lies in [i-7, i+7] (the clamps of
        // lib.rs:95-110 only move x towards the interior) and i0+1 <= i+8; rows likewise within [j-3, j+4].  The
        // NaN-keeping max makes a NaN displacement fail the test.
        float du[4], dv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { du[c] = fmul(dt0, uc[c]); dv[c] = fmul(dt1, vc[c]); }
        const float mx = max_keep_nan(max_keep_nan(fabsf(du[0]), fabsf(du[1])), max_keep_nan(fabsf(du[2]), fabsf(du[3])));
        const float my = max_keep_nan(max_keep_nan(fabsf(dv[0]), fabsf(dv[1])), max_keep_nan(fabsf(dv[2]), fabsf(dv[3])));
        const bool near = (mx <= (float)(ADV_MX - 1)) && (my <= (float)(ADV_MY - 1));
        Trace t[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // backtrace(), kernels_basic.cuh, with i and j already converted
            float x = fsub(fi[c], du[c]);
            float y = fsub(fj, dv[c]);
            x = min_keep_nan(max_keep_nan(x, 0.5f), xmax);
            y = min_keep_nan(max_keep_nan(y, 0.5f), ymax);
            t[c].i0 = (int)__float2uint_rz(x);
            t[c].j0 = (int)__float2uint_rz(y);
            t[c].s1 = fsub(x, (float)t[c].i0);
            t[c].t1 = fsub(y, (float)t[c].j0);
        }
        float r[NF][4];
        if (near) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float* const p = win + (t[c].j0 * ADV_WW + t[c].i0 - wbase);
                const float t0 = fsub(1.0f, t[c].t1), s0 = fsub(1.0f, t[c].s1);
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const float a00 = p[f * ADV_FIELD_FLOATS], a10 = p[f * ADV_FIELD_FLOATS + 1];
                    const float a01 = p[f * ADV_FIELD_FLOATS + ADV_WW], a11 = p[f * ADV_FIELD_FLOATS + ADV_WW + 1];
                    // mix(mix(a00, a01, t1), mix(a10, a11, t1), s1), lib.rs:80-82,118-121
                    const float m0 = fadd(fmul(t0, a00), fmul(t[c].t1, a01));
                    const float m1 = fadd(fmul(t0, a10), fmul(t[c].t1, a11));
                    r[f][c] = fadd(fmul(s0, m0), fmul(t[c].s1, m1));
                }
            }
        } else {
            // A back-trace may leave the window (CFL beyond the margins): this thread takes its four cells from global
            // memory -- all loads first, then the arithmetic; taps are always inside the field (lib.rs:95-110), also
            // for the columns outside the interior, whose results are never stored.
            size_t q[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                q[c] = at(g, t[c].i0, t[c].j0);
                // a slab holds only its own rows: taps elsewhere are pass 2's business, load something harmless instead
                if (PASS == 1 && !((t[c].j0 >= tab.my_lo) && (t[c].j0 + 1 < tab.my_hi))) q[c] = (size_t)j * g.pitch;
            }
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                float a00[4], a10[4], a01[4], a11[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float* const p = A.a.d0[f] + q[c];
                    a00[c] = __ldg(p); a10[c] = __ldg(p + 1); a01[c] = __ldg(p + g.pitch); a11[c] = __ldg(p + g.pitch + 1);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    r[f][c] = mixf(mixf(a00[c], a01[c], t[c].t1), mixf(a10[c], a11[c], t[c].t1), t[c].s1);
            }
        }
        bool own[4] = {true, true, true, true};
        // A block at least a window margin away from both slab cuts whose warp took the window path throughout has all
        // its taps in owned rows ([j-3, j+4] around rows of the block): no vote needed (all but two block rows per rank).
        if (PASS == 1 && !(deep && __all_sync(0xffffffffu, near))) {
            // k_advect4's pass 2 decides per thread = per quad of aligned columns [4q, 4q+4) (all of them, border and
            // padding columns of the field included) whether the quad's taps are all in rows this rank owns.
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const bool mine = (t[c].j0 >= tab.my_lo) && (t[c].j0 + 1 < tab.my_hi);
                const unsigned quad = (__ballot_sync(0xffffffffu, mine) >> (lane & 28)) & 0xfu;
                const bool counted = ((i_first + 32 * c) & ~3) < g.w;  // k_advect4 returns early for gx0 >= w
                own[c] = (quad == 0xfu);
                if (!own[c] && counted && (lane & 3) == 0) {  // queue the quad's block once
                    const int bid = ((jl - A.ly0) >> 3) * tab.grid_x + (int)blockIdx.x;
                    if (atomicExch(&tab.block_flag[bid], 1) == 0) tab.block_list[1 + atomicAdd(&tab.block_list[0], 1)] = bid;
                }
            }
        }
        if (!edge && PASS == 0) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                float* const orow = A.a.d[f] + krow;
#pragma unroll
                for (int c = 0; c < 4; ++c) orow[32 * c] = r[f][c];
            }
        } else if (!edge) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                float* const orow = A.a.d[f] + krow;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (own[c]) orow[32 * c] = r[f][c];
            }
        } else {
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (col_ok[c] && own[c]) store_with_bnd_rows(A.a.d[f], g, A.a.b[f], i_first + 32 * c, jl, r[f][c]);
        }
    }
}

// Tensor maps of the gathered fields + launch.  Returns a cudaError_t as int, or -4 when no tensor map can be made.
template <int NF, int UV, int PASS>
static int advect_tile_launch(const AdvectArgs& a, const float* u, const float* v, const Geom& g, float dt0, float dt1,
                              float xmax, float ymax, int ly0, int ly1, const SlabTable& tab, cudaStream_t st) {
    AdvectTileArgs A;
    for (int f = 0; f < NF; ++f)  // the windows are local rows; a.d0[] is indexed by global row
        if (int r = field_tensor_map(a.d0[f] + (size_t)g.gy0 * g.pitch, g.pitch, g.h, ADV_WW, ADV_WH, &A.tm[f])) return r;
    if (UV == ADV_UV_TMA) {
        if (int r = field_tensor_map(u, g.pitch, g.h, ADV_TW, ADV_TH, &A.tm_u)) return r;
        if (int r = field_tensor_map(v, g.pitch, g.h, ADV_TW, ADV_TH, &A.tm_v)) return r;
    }
    A.a = a; A.u = u; A.v = v; A.g = g; A.dt0 = dt0; A.dt1 = dt1; A.xmax = xmax; A.ymax = ymax;
    A.ly0 = ly0; A.ly1 = ly1;
    constexpr size_t smem = advect_tile_smem_bytes<NF, UV>();
    if (smem > 48 * 1024) {
        static std::atomic<uint64_t> attr_devices{0};
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
        if (!((attr_devices.load(std::memory_order_acquire) >> dev) & 1)) {
            const cudaError_t e = cudaFuncSetAttribute(k_advect_tile<NF, UV, PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr_devices.fetch_or(1ull << dev, std::memory_order_release);
        }
    }
    const dim3 grid((g.w + ADV_TW - 1) / ADV_TW, (ly1 - ly0 + ADV_TH - 1) / ADV_TH);
    return (int)launch_pdl(true, k_advect_tile<NF, UV, PASS>, grid, dim3(256), smem, st, A, tab);
}

}  // namespace fruid
