#define QMLB_T double
#define QMLB_LAUNCH_TILE launch_tile_f64
#define QMLB_TILE_SET_SMEM tile_set_smem_f64
#include "qmlb_tile_inst.cuh"
