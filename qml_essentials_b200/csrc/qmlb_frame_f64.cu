#define QMLB_T double
#define QMLB_LAUNCH_FRAME launch_frame_f64
#include "qmlb_frame_inst.cuh"
