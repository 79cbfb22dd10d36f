#define QMLB_T float
#define QMLB_LAUNCH_FRAME launch_frame_f32
#include "qmlb_frame_inst.cuh"
