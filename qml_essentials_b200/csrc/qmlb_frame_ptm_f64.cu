#define QMLB_T double
#define QMLB_LAUNCH_FRAME_PTM launch_frame_ptm_f64
#include "qmlb_frame_ptm_inst.cuh"
