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
//   * a cell whose back-trace leaves the window (|displacement| larger than the margin: CFL > 4) takes the global
//     gather of kernels_basic.cuh instead -- per cell, so results never depend on which path ran;
//   * velocity self-advection (lib.rs:204-205: the gathered fields ARE the advecting velocity) reads the cell's own
//     velocity from the window too (SELF), i.e. u0, v0 are read once;
//   * set_bnd (lib.rs:125) is folded into the stores; only blocks that touch the domain border pay for its tests.
// Arithmetic is the scalar code of kernels_basic.cuh (backtrace(), mixf()): bit-identical to k_advect / k_advect4.
#pragma once
#include "jacobi_tile.cuh"
#include "kernels_basic.cuh"

namespace fruid {

constexpr int ADV_TW = 128, ADV_TH = 32;          // cells produced per block
constexpr int ADV_MX = 8, ADV_MY = 4;             // window margins (columns: a multiple of 4 keeps rows 16-byte aligned)
constexpr int ADV_WW = ADV_TW + 2 * ADV_MX;       // 144 floats per window row (576 B)
constexpr int ADV_WH = ADV_TH + 2 * ADV_MY;       // 40 rows
constexpr int ADV_FIELD_FLOATS = ADV_WW * ADV_WH; // 5760 floats = 23040 B per field

struct AdvectTileArgs {
    CUtensorMap tm[ADVECT_MAX_FIELDS];  // gathered fields d0[f]: dims (pitch, h), box (ADV_WW, ADV_WH)
    AdvectArgs a;
    const float* u;
    const float* v;
    Geom g;
    float dt0, dt1, xmax, ymax;
};

template <int NF>
constexpr size_t advect_tile_smem_bytes() { return (size_t)NF * ADV_FIELD_FLOATS * sizeof(float) + 16; }

template <int NF, bool SELF>
__global__ void __launch_bounds__(256) k_advect_tile(const __grid_constant__ AdvectTileArgs A) {
    pdl_enter();
    extern __shared__ __align__(128) unsigned char adv_smem[];
    float* const win = reinterpret_cast<float*>(adv_smem);
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(adv_smem + (size_t)NF * ADV_FIELD_FLOATS * sizeof(float));
    const Geom& g = A.g;
    const int bx0 = blockIdx.x * ADV_TW;           // first column of the block (column 0 is the border)
    const int by0 = 1 + blockIdx.y * ADV_TH;       // first (interior) row of the block
    const int wx0 = bx0 - ADV_MX, wy0 = by0 - ADV_MY;  // window origin
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        mbar_expect_tx(mbar, (uint32_t)(NF * ADV_FIELD_FLOATS * sizeof(float)));
#pragma unroll
        for (int f = 0; f < NF; ++f) tma_load_2d(win + f * ADV_FIELD_FLOATS, &A.tm[f], wx0, wy0, mbar);
    }
    __syncthreads();  // the barrier is initialised before anybody waits on it
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    // set_bnd tests only where the block touches the border (block-uniform)
    const bool edge = (bx0 == 0) || (bx0 + ADV_TW >= g.nx) || (by0 == 1) || (by0 + ADV_TH > g.ny);
    mbar_wait(mbar, 0);
    dbg_tail_delay(wi);
#pragma unroll 1
    for (int rr = 0; rr < 4; ++rr) {
        const int j = by0 + wi * 4 + rr;
        if (j > g.ny) break;  // warp-uniform
        const float* const wrow = win + (j - wy0) * ADV_WW + (lane + ADV_MX);
        float uc[4], vc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = bx0 + lane + 32 * c;
            if (SELF) {
                uc[c] = wrow[32 * c];
                vc[c] = wrow[32 * c + ADV_FIELD_FLOATS];
            } else {
                const bool in = (i >= 1 && i <= g.nx);
                const size_t k = at(g, i, j);
                uc[c] = in ? __ldg(A.u + k) : 0.f;
                vc[c] = in ? __ldg(A.v + k) : 0.f;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = bx0 + lane + 32 * c;
            if (i < 1 || i > g.nx) continue;
            const Trace t = backtrace(i, j, uc[c], vc[c], A.dt0, A.dt1, A.xmax, A.ymax);
            const unsigned wx = (unsigned)(t.i0 - wx0), wy = (unsigned)(t.j0 - wy0);
            float r[NF];
            if (wx <= (unsigned)(ADV_WW - 2) && wy <= (unsigned)(ADV_WH - 2)) {  // taps (wx, wx+1) x (wy, wy+1) inside
                const float* p = win + wy * ADV_WW + wx;
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const float a00 = p[f * ADV_FIELD_FLOATS], a10 = p[f * ADV_FIELD_FLOATS + 1];
                    const float a01 = p[f * ADV_FIELD_FLOATS + ADV_WW], a11 = p[f * ADV_FIELD_FLOATS + ADV_WW + 1];
                    r[f] = mixf(mixf(a00, a01, t.t1), mixf(a10, a11, t.t1), t.s1);
                }
            } else {
#pragma unroll
                for (int f = 0; f < NF; ++f) r[f] = gather(A.a.d0[f], g, t);
            }
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                if (edge) store_with_bnd(A.a.d[f], g, A.a.b[f], i, j, r[f]);
                else A.a.d[f][at(g, i, j)] = r[f];
            }
        }
    }
}

// Tensor maps of the gathered fields + launch.  Returns a cudaError_t as int, or -4 when no tensor map can be made.
template <int NF, bool SELF>
static int advect_tile_launch(const AdvectArgs& a, const float* u, const float* v, const Geom& g, float dt0, float dt1,
                              float xmax, float ymax, cudaStream_t st) {
    AdvectTileArgs A;
    for (int f = 0; f < NF; ++f)
        if (int r = field_tensor_map(a.d0[f], g.pitch, g.h, ADV_WW, ADV_WH, &A.tm[f])) return r;
    A.a = a; A.u = u; A.v = v; A.g = g; A.dt0 = dt0; A.dt1 = dt1; A.xmax = xmax; A.ymax = ymax;
    constexpr size_t smem = advect_tile_smem_bytes<NF>();
    if (smem > 48 * 1024) {
        static std::atomic<uint64_t> attr_devices{0};
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
        if (!((attr_devices.load(std::memory_order_acquire) >> dev) & 1)) {
            const cudaError_t e = cudaFuncSetAttribute(k_advect_tile<NF, SELF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr_devices.fetch_or(1ull << dev, std::memory_order_release);
        }
    }
    const dim3 grid((g.w + ADV_TW - 1) / ADV_TW, (g.ny + ADV_TH - 1) / ADV_TH);
    return (int)launch_pdl(true, k_advect_tile<NF, SELF>, grid, dim3(256), smem, st, A);
}

}  // namespace fruid
