#define QMLB_T double
#define QMLB_LAUNCH_REG launch_reg_f64
#include "qmlb_reg_inst.cuh"
