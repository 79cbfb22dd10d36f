#define QMLB_T float
#define QMLB_LAUNCH_REG launch_reg_f32
#include "qmlb_reg_inst.cuh"
