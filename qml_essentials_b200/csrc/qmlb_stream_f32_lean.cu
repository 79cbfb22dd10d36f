#define QMLB_T float
#define QMLB_STREAM_R 4
#define QMLB_STREAM_HEAVY 0
#define QMLB_LAUNCH_STREAM launch_stream_f32_lean
#define QMLB_LAUNCH_STREAM_MATS launch_stream_mats_f32
#include "qmlb_stream_inst.cuh"
