#define QMLB_T double
#define QMLB_LAUNCH_FSTREAM launch_fstream_f64
#include "qmlb_fstream_inst.cuh"
