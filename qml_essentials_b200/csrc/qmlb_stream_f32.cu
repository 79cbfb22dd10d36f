#define QMLB_T float
#define QMLB_STREAM_R 4
#define QMLB_LAUNCH_STREAM launch_stream_f32
#define QMLB_STREAM_SET_SMEM stream_set_smem_f32
#define QMLB_LAUNCH_STREAM_MATS launch_stream_mats_f32
#include "qmlb_stream_inst.cuh"
