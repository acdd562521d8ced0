// kern_stub_any.cu -- tuning builds only (tools/ab_multi.sh): stands in for every transform size that is NOT kept with
// -DKEEP_<M>, for both the classic and the warp-specialised chain, so that an A/B library with one or two sizes compiles in
// well under a minute.  Never part of libbtkb200.so.
//   AB_SRCS="kern_m256.cu kern_ws_m256.cu kern_stub_any.cu kern_misc.cu kern_cov_tc.cu kern_postfilter.cu kern_design.cu capi.cu"
//   tools/ab_multi.sh a "-DKEEP_256 -DFOO" b "-DKEEP_256"
#include "launch.h"
namespace btk {
#define STUB(MM)                                                                                                   \
  cudaError_t launch_chain_m##MM(int, const ChainParams&, int, cudaStream_t) { return cudaErrorInvalidValue; }       \
  cudaError_t launch_analysis_m##MM(int, const AnalysisParams&, int, cudaStream_t) { return cudaErrorInvalidValue; } \
  cudaError_t launch_synthesis_m##MM(int, const SynthesisParams&, int, cudaStream_t) { return cudaErrorInvalidValue; } \
  int fb_smem_bytes_m##MM(int, int) { return -1; }                                                                 \
  int chain_frames_per_iter_m##MM(int, int) { return -1; }                                                         \
  cudaError_t launch_chain_ws_m##MM(int, const ChainParams&, int, cudaStream_t) { return cudaErrorInvalidValue; }    \
  int chain_ws_frames_per_iter_m##MM(int, int) { return -1; }                                                      \
  bool chain_ws_cluster_ok_m##MM(int, int, int) { return false; }
#ifndef KEEP_64
STUB(64)
#endif
#ifndef KEEP_128
STUB(128)
#endif
#ifndef KEEP_256
STUB(256)
#endif
#ifndef KEEP_512
STUB(512)
#endif
#ifndef KEEP_1024
STUB(1024)
#endif
}  // namespace btk
