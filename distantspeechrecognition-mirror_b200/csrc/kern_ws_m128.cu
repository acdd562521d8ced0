// kern_ws_m128.cu -- warp-specialised fused chain for M = 128 (all decimation factors R = 1, 2, 4, 8).
#include "kern_ws.cuh"
BTK_DEFINE_WS_LAUNCHERS(128)
