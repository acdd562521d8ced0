// kern_stub1024.cu -- tuning builds only (tools/ab_multi.sh): like kern_stub.cu but leaves M = 1024 to kern_m1024.cu.
#include "launch.h"
namespace btk {
#define STUB(MM)                                                                                                   \
  cudaError_t launch_chain_m##MM(int, const ChainParams&, int, cudaStream_t) { return cudaErrorInvalidValue; }       \
  cudaError_t launch_analysis_m##MM(int, const AnalysisParams&, int, cudaStream_t) { return cudaErrorInvalidValue; } \
  cudaError_t launch_synthesis_m##MM(int, const SynthesisParams&, int, cudaStream_t) { return cudaErrorInvalidValue; } \
  int fb_smem_bytes_m##MM(int, int) { return -1; }                                                                 \
  int chain_frames_per_iter_m##MM(int, int) { return -1; }
STUB(64) STUB(128) STUB(256) STUB(512)
}  // namespace btk
