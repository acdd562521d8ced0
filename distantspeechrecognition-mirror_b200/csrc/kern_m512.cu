// kern_m512.cu -- filter-bank kernels for M = 512 (all decimation factors R = 1, 2, 4, 8).
#include "kern_fb.cuh"
BTK_DEFINE_M_LAUNCHERS(512)
