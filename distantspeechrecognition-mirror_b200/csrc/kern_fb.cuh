// kern_fb.cuh -- __global__ wrappers of the tile programs + per-M launch dispatch.
// Included by kern_m<M>.cu with BTK_KERN_M defined, so every transform size is its own translation unit
// (parallel nvcc, bounded compile time).
//
// Every kernel exists in a generic form (MT = 0: prototype length factor m read at run time) and in
// specialised forms with m as a compile-time constant (MT = 2, 4: the factors of the reference's shipped
// prototypes, btk/examples/prototypes/Nyquist) whose polyphase loops unroll completely.
#pragma once

#include "launch.h"

namespace btk {

template <int M> struct DevCtx {
  ChainThreadState<M> ts;
  template <class F> __device__ __forceinline__ void par(F f) { f((int)threadIdx.x, ts); }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void syncwarp() { __syncwarp(); }
};

// two CTAs per SM whenever two of them fit the 227 KB of shared memory
template <int M, int R, int MT> struct KernCfg {
  static constexpr int smem = chain_smem_layout<M, R>(MT > 0 ? MT : 4).total;
  static constexpr int MINB = (M <= 256 && smem <= 113 * 1024) ? 2 : 1;
};

template <int M, int R, int MT>
__global__ void __launch_bounds__(ChainCfg<M, R, MT>::NT, KernCfg<M, R, MT>::MINB) btk_chain_kernel(const ChainParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  DevCtx<M> ctx;
  chain_tile<M, R, MT>(ctx, p, smem, (int)blockIdx.x);
}

template <int M, int R, int MT>
__global__ void __launch_bounds__(ChainCfg<M, R, MT>::NT, KernCfg<M, R, MT>::MINB) btk_analysis_kernel(const AnalysisParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  DevCtx<M> ctx;
  analysis_tile<M, R, MT>(ctx, p, smem, (int)blockIdx.x);
}

template <int M, int R, int MT>
__global__ void __launch_bounds__(ChainCfg<M, R, MT>::NT, KernCfg<M, R, MT>::MINB) btk_synthesis_kernel(const SynthesisParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  DevCtx<M> ctx;
  synthesis_tile<M, R, MT>(ctx, p, smem, (int)blockIdx.x);
}

template <int M, int R, class Params, class Kern>
static cudaError_t launch_one(Kern kern, const Params& p, int m, int n_work, cudaStream_t st) {
  const ChainSmem L = chain_smem_layout<M, R>(m);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  if (e != cudaSuccess) return e;
  kern<<<n_work, ChainCfg<M, R, 0>::NT, L.total, st>>>(p);
  return cudaGetLastError();
}

// compile-time m where the unrolled polyphase stays within the register budget
template <int M, int R, int MT> struct FastOk { static constexpr bool value = MT * R <= 16; };

template <int M, int R>
static cudaError_t launch_chain_r(const ChainParams& p, int n_work, cudaStream_t st) {
  if (p.m == 2 && FastOk<M, R, 2>::value) return launch_one<M, R>(btk_chain_kernel<M, R, FastOk<M, R, 2>::value ? 2 : 0>, p, p.m, n_work, st);
  if (p.m == 4 && FastOk<M, R, 4>::value) return launch_one<M, R>(btk_chain_kernel<M, R, FastOk<M, R, 4>::value ? 4 : 0>, p, p.m, n_work, st);
  return launch_one<M, R>(btk_chain_kernel<M, R, 0>, p, p.m, n_work, st);
}
template <int M, int R>
static cudaError_t launch_analysis_r(const AnalysisParams& p, int n_work, cudaStream_t st) {
  if (p.m == 2 && FastOk<M, R, 2>::value) return launch_one<M, R>(btk_analysis_kernel<M, R, FastOk<M, R, 2>::value ? 2 : 0>, p, p.m, n_work, st);
  if (p.m == 4 && FastOk<M, R, 4>::value) return launch_one<M, R>(btk_analysis_kernel<M, R, FastOk<M, R, 4>::value ? 4 : 0>, p, p.m, n_work, st);
  return launch_one<M, R>(btk_analysis_kernel<M, R, 0>, p, p.m, n_work, st);
}
template <int M, int R>
static cudaError_t launch_synthesis_r(const SynthesisParams& p, int n_work, cudaStream_t st) {
  if (p.m == 2 && FastOk<M, R, 2>::value) return launch_one<M, R>(btk_synthesis_kernel<M, R, FastOk<M, R, 2>::value ? 2 : 0>, p, p.m, n_work, st);
  if (p.m == 4 && FastOk<M, R, 4>::value) return launch_one<M, R>(btk_synthesis_kernel<M, R, FastOk<M, R, 4>::value ? 4 : 0>, p, p.m, n_work, st);
  return launch_one<M, R>(btk_synthesis_kernel<M, R, 0>, p, p.m, n_work, st);
}

}  // namespace btk

// Defines launch_*_m<M>(R, ...) for the including translation unit.
#define BTK_DEFINE_M_LAUNCHERS(MM)                                                                                 \
  namespace btk {                                                                                                  \
  cudaError_t launch_chain_m##MM(int R, const ChainParams& p, int n_work, cudaStream_t st) {                       \
    switch (R) {                                                                                                   \
      case 1: return launch_chain_r<MM, 1>(p, n_work, st);                                                         \
      case 2: return launch_chain_r<MM, 2>(p, n_work, st);                                                         \
      case 4: return launch_chain_r<MM, 4>(p, n_work, st);                                                         \
      case 8: return launch_chain_r<MM, 8>(p, n_work, st);                                                         \
    }                                                                                                              \
    return cudaErrorInvalidValue;                                                                                  \
  }                                                                                                                \
  cudaError_t launch_analysis_m##MM(int R, const AnalysisParams& p, int n_work, cudaStream_t st) {                 \
    switch (R) {                                                                                                   \
      case 1: return launch_analysis_r<MM, 1>(p, n_work, st);                                                      \
      case 2: return launch_analysis_r<MM, 2>(p, n_work, st);                                                      \
      case 4: return launch_analysis_r<MM, 4>(p, n_work, st);                                                      \
      case 8: return launch_analysis_r<MM, 8>(p, n_work, st);                                                      \
    }                                                                                                              \
    return cudaErrorInvalidValue;                                                                                  \
  }                                                                                                                \
  cudaError_t launch_synthesis_m##MM(int R, const SynthesisParams& p, int n_work, cudaStream_t st) {               \
    switch (R) {                                                                                                   \
      case 1: return launch_synthesis_r<MM, 1>(p, n_work, st);                                                     \
      case 2: return launch_synthesis_r<MM, 2>(p, n_work, st);                                                     \
      case 4: return launch_synthesis_r<MM, 4>(p, n_work, st);                                                     \
      case 8: return launch_synthesis_r<MM, 8>(p, n_work, st);                                                     \
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
