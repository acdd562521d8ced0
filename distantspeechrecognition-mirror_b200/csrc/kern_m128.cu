// kern_m128.cu -- filter-bank kernels for M = 128 (all decimation factors R = 1, 2, 4, 8).
#include "kern_fb.cuh"
BTK_DEFINE_M_LAUNCHERS(128)
