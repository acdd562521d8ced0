// kern_fb.cuh -- __global__ wrappers of the tile programs + per-M launch dispatch.
// Included by kern_m<M>.cu with BTK_KERN_M defined, so every transform size is its own translation unit
// (parallel nvcc, bounded compile time).
#pragma once

#include "launch.h"

namespace btk {

template <int M> struct DevCtx {
  ChainThreadState<M> ts;
  template <class F> __device__ __forceinline__ void par(F f) { f((int)threadIdx.x, ts); }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void syncwarp() { __syncwarp(); }
};

template <int M, int R>
__global__ void __launch_bounds__(ChainCfg<M, R>::NT, (M <= 256 ? 2 : 1)) btk_chain_kernel(const ChainParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  DevCtx<M> ctx;
  chain_tile<M, R>(ctx, p, smem, (int)blockIdx.x);
}

template <int M, int R>
__global__ void __launch_bounds__(ChainCfg<M, R>::NT, (M <= 256 ? 2 : 1)) btk_analysis_kernel(const AnalysisParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  DevCtx<M> ctx;
  analysis_tile<M, R>(ctx, p, smem, (int)blockIdx.x);
}

template <int M, int R>
__global__ void __launch_bounds__(ChainCfg<M, R>::NT, (M <= 256 ? 2 : 1)) btk_synthesis_kernel(const SynthesisParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  DevCtx<M> ctx;
  synthesis_tile<M, R>(ctx, p, smem, (int)blockIdx.x);
}

template <int M, int R, class Params, class Kern>
static cudaError_t launch_one(Kern kern, const Params& p, int m, int n_work, cudaStream_t st) {
  const ChainSmem L = chain_smem_layout<M, R>(m);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  if (e != cudaSuccess) return e;
  kern<<<n_work, ChainCfg<M, R>::NT, L.total, st>>>(p);
  return cudaGetLastError();
}

}  // namespace btk

// Defines launch_*_m<M>(R, ...) for the including translation unit.
#define BTK_DEFINE_M_LAUNCHERS(MM)                                                                                 \
  namespace btk {                                                                                                  \
  cudaError_t launch_chain_m##MM(int R, const ChainParams& p, int n_work, cudaStream_t st) {                       \
    switch (R) {                                                                                                   \
      case 1: return launch_one<MM, 1>(btk_chain_kernel<MM, 1>, p, p.m, n_work, st);                               \
      case 2: return launch_one<MM, 2>(btk_chain_kernel<MM, 2>, p, p.m, n_work, st);                               \
      case 4: return launch_one<MM, 4>(btk_chain_kernel<MM, 4>, p, p.m, n_work, st);                               \
      case 8: return launch_one<MM, 8>(btk_chain_kernel<MM, 8>, p, p.m, n_work, st);                               \
    }                                                                                                              \
    return cudaErrorInvalidValue;                                                                                  \
  }                                                                                                                \
  cudaError_t launch_analysis_m##MM(int R, const AnalysisParams& p, int n_work, cudaStream_t st) {                 \
    switch (R) {                                                                                                   \
      case 1: return launch_one<MM, 1>(btk_analysis_kernel<MM, 1>, p, p.m, n_work, st);                            \
      case 2: return launch_one<MM, 2>(btk_analysis_kernel<MM, 2>, p, p.m, n_work, st);                            \
      case 4: return launch_one<MM, 4>(btk_analysis_kernel<MM, 4>, p, p.m, n_work, st);                            \
      case 8: return launch_one<MM, 8>(btk_analysis_kernel<MM, 8>, p, p.m, n_work, st);                            \
    }                                                                                                              \
    return cudaErrorInvalidValue;                                                                                  \
  }                                                                                                                \
  cudaError_t launch_synthesis_m##MM(int R, const SynthesisParams& p, int n_work, cudaStream_t st) {               \
    switch (R) {                                                                                                   \
      case 1: return launch_one<MM, 1>(btk_synthesis_kernel<MM, 1>, p, p.m, n_work, st);                           \
      case 2: return launch_one<MM, 2>(btk_synthesis_kernel<MM, 2>, p, p.m, n_work, st);                           \
      case 4: return launch_one<MM, 4>(btk_synthesis_kernel<MM, 4>, p, p.m, n_work, st);                           \
      case 8: return launch_one<MM, 8>(btk_synthesis_kernel<MM, 8>, p, p.m, n_work, st);                           \
    }                                                                                                              \
    return cudaErrorInvalidValue;                                                                                  \
  }                                                                                                                \
  int fb_smem_bytes_m##MM(int R, int m) {                                                                          \
    switch (R) {                                                                                                   \
      case 1: return chain_smem_layout<MM, 1>(m).total;                                                            \
      case 2: return chain_smem_layout<MM, 2>(m).total;                                                            \
      case 4: return chain_smem_layout<MM, 4>(m).total;                                                            \
      case 8: return chain_smem_layout<MM, 8>(m).total;                                                            \
    }                                                                                                              \
    return -1;                                                                                                     \
  }                                                                                                                \
  }
