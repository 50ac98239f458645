// Tile shapes of the fused Jacobi kernel, group a (the default 128x128 shape and its neighbours).
#include "tile_inst.cuh"
#define SHAPES_A(X) X(8, 16) X(10, 16) X(8, 12)
FRUID_TILE_INST_UNIT(jacobi_tile_launch_inst_a, SHAPES_A)
