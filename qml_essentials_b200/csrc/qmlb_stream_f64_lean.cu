#define QMLB_T double
#define QMLB_STREAM_R 4
#define QMLB_STREAM_HEAVY 0
#define QMLB_LAUNCH_STREAM launch_stream_f64_lean
#define QMLB_LAUNCH_STREAM_MATS launch_stream_mats_f64
#include "qmlb_stream_inst.cuh"
