// tile_inst.cuh -- instantiates k_jacobi_tile for a group of tile shapes and defines that group's launcher.
// Included by tile_inst_{a,b,c,d}.cu (one translation unit each: the library builds in parallel).
#pragma once
#include <atomic>
#include "jacobi_tile_kernel.cuh"

namespace fruid {

template <int R, int NW, int MODE>
static int tile_launch_mode(const TileArgs& A, int num_sms, bool pdl, cudaStream_t st) {
    constexpr size_t smem = jacobi_tile_smem_bytes<R, NW>();
    static std::atomic<uint64_t> attr_devices{0};  // per instantiation: devices whose function attribute is set
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    if (!((attr_devices.load(std::memory_order_acquire) >> dev) & 1)) {
        const cudaError_t e = cudaFuncSetAttribute(k_jacobi_tile<R, NW, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_devices.fetch_or(1ull << dev, std::memory_order_release);
    }
    const int ntiles = A.ntx * A.nty;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ntiles < num_sms ? ntiles : num_sms);
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, k_jacobi_tile<R, NW, MODE>, A);
}

template <int R, int NW>
static int tile_launch_cfg(const TileArgs& A, int mode, int num_sms, bool pdl, cudaStream_t st) {
    if (mode == MODE_GENERAL) return tile_launch_mode<R, NW, MODE_GENERAL>(A, num_sms, pdl, st);
    if (mode == MODE_POW2) return tile_launch_mode<R, NW, MODE_POW2>(A, num_sms, pdl, st);
    if (mode == MODE_ZEROA) return tile_launch_mode<R, NW, MODE_ZEROA>(A, num_sms, pdl, st);
    if (mode == MODE_POISSON) return tile_launch_mode<R, NW, MODE_POISSON>(A, num_sms, pdl, st);
    if (mode == MODE_RECIP) return tile_launch_mode<R, NW, MODE_RECIP>(A, num_sms, pdl, st);
    if (mode == MODE_RECIP3) return tile_launch_mode<R, NW, MODE_RECIP3>(A, num_sms, pdl, st);
    return tile_launch_mode<R, NW, MODE_ONE>(A, num_sms, pdl, st);
}

}  // namespace fruid

#define FRUID_TILE_INST_CASE(r, nw) if (R == r && NW == nw) return tile_launch_cfg<r, nw>(A, mode, num_sms, pdl, st);
#define FRUID_TILE_INST_UNIT(name, SHAPES)                                                                              \
    namespace fruid {                                                                                                   \
    int name(int R, int NW, int mode, const TileArgs& A, int num_sms, bool pdl, cudaStream_t st) {                      \
        SHAPES(FRUID_TILE_INST_CASE)                                                                                    \
        return -1;                                                                                                      \
    }                                                                                                                   \
    }
