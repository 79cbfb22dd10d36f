#define QMLB_T float
#define QMLB_LAUNCH_FSTREAM launch_fstream_f32
#include "qmlb_fstream_inst.cuh"
