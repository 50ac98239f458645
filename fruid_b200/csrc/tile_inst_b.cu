// Tile shapes of the fused Jacobi kernel, group b.
#include "tile_inst.cuh"
#define SHAPES_B(X) X(6, 16) X(6, 20) X(7, 20)
FRUID_TILE_INST_UNIT(jacobi_tile_launch_inst_b, SHAPES_B)
