// jacobi_tile.cuh -- T fused Jacobi sweeps per launch (placeholder until the
// register-tile kernel lands; the default is one sweep per launch).
#pragma once
#include "common.cuh"
namespace fruid {
static inline int jacobi_tile_default_fused(const Geom&, int) { return 1; }
static inline int jacobi_tile_launch(float*, const float*, const float*, const Geom&, float, float, int, int,
                                     cudaStream_t) {
    return -7;
}
}  // namespace fruid
