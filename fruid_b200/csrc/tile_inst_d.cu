// Tile shapes of the fused Jacobi kernel, group d (the small-grid shapes).
#include "tile_inst.cuh"
#define SHAPES_D(X) X(2, 32) X(2, 16) X(3, 16)
FRUID_TILE_INST_UNIT(jacobi_tile_launch_inst_d, SHAPES_D)
