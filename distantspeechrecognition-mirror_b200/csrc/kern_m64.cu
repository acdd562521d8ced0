// kern_m64.cu -- filter-bank kernels for M = 64 (all decimation factors R = 1, 2, 4, 8).
#include "kern_fb.cuh"
BTK_DEFINE_M_LAUNCHERS(64)
