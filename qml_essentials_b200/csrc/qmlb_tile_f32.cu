#define QMLB_T float
#define QMLB_LAUNCH_TILE launch_tile_f32
#define QMLB_TILE_SET_SMEM tile_set_smem_f32
#include "qmlb_tile_inst.cuh"
