#define QMLB_T float
#define QMLB_LAUNCH_FRAME_PTM launch_frame_ptm_f32
#include "qmlb_frame_ptm_inst.cuh"
