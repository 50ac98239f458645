// Tile shapes of the fused Jacobi kernel, group c.
#include "tile_inst.cuh"
#define SHAPES_C(X) X(5, 24) X(4, 24) X(4, 16) X(4, 8)
FRUID_TILE_INST_UNIT(jacobi_tile_launch_inst_c, SHAPES_C)
