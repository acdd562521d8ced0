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

template <int M, int PP> struct DevCtx {
  ChainThreadState<M, PP> ts;
  template <class F> __device__ __forceinline__ void par(F f) { f((int)threadIdx.x, ts); }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void syncwarp() { __syncwarp(); }
};

#define BTK_MAX_SMEM (227 * 1024)

// two CTAs per SM whenever two of them fit the 227 KB of shared memory
template <int M, int R, int MT, int PP> struct KernCfg {
  static constexpr int smem = chain_smem_layout<M, R, PP>(MT > 0 ? MT : 4).total;
  static constexpr int MINB = (smem <= 112 * 1024) ? 2 : 1;
};

// Frame pairs per warp of the fused chain: two for M <= 256 (four-warp CTAs, two per SM, ~230 registers per thread; taps /
// samples / weights / twiddles loaded once for four frames) whenever the larger window fits shared memory; one for the
// 32-values-per-lane transforms of M = 512 / 1024.
#ifndef BTK_PP_MAX_M
#define BTK_PP_MAX_M 256   // M = 512 with two pairs per warp measured equal (16 ch) or 44 % slower (64 ch): tools/ab_run2.sh
#endif
template <int M, int R> static int chain_pp(int m) {
  return (M <= BTK_PP_MAX_M && chain_smem_layout<M, R, 2>(m).total <= BTK_MAX_SMEM) ? 2 : 1;
}

template <int M, int R, int MT, int PP>
__global__ void __launch_bounds__(ChainCfg<M, R, MT, PP>::NT, KernCfg<M, R, MT, PP>::MINB) btk_chain_kernel(const ChainParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  DevCtx<M, PP> ctx;
  chain_tile<M, R, MT, PP>(ctx, p, smem, (int)blockIdx.x);
}

template <int M, int R, int MT>
__global__ void __launch_bounds__(ChainCfg<M, R, MT>::NT, KernCfg<M, R, MT, 1>::MINB) btk_analysis_kernel(const AnalysisParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  DevCtx<M, 1> ctx;
  analysis_tile<M, R, MT>(ctx, p, smem, (int)blockIdx.x);
}

template <int M, int R, int MT>
__global__ void __launch_bounds__(ChainCfg<M, R, MT>::NT, KernCfg<M, R, MT, 1>::MINB) btk_synthesis_kernel(const SynthesisParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  DevCtx<M, 1> ctx;
  synthesis_tile<M, R, MT>(ctx, p, smem, (int)blockIdx.x);
}

// only the fused chain carries the hint; the staged kernels keep their natural occupancy
inline bool one_cta_per_sm(const ChainParams& p) { return p.one_cta != 0; }
inline bool one_cta_per_sm(const AnalysisParams&) { return false; }
inline bool one_cta_per_sm(const SynthesisParams&) { return false; }

template <int M, int R, int PP, class Params, class Kern>
static cudaError_t launch_one(Kern kern, const Params& p, int m, int n_work, cudaStream_t st) {
  const ChainSmem L = chain_smem_layout<M, R, PP>(m);
  // one_cta_per_sm(p): the launch asks for more than half of the SM's shared memory, so that only one CTA is resident
  // (see chain_one_cta_per_sm below: many-channel inputs whose concurrent windows would not fit L2 otherwise)
  int smem = L.total;
  if (one_cta_per_sm(p) && smem < 116 * 1024) smem = 116 * 1024;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  kern<<<n_work, ChainCfg<M, R, 0, PP>::NT, smem, st>>>(p);
  return cudaGetLastError();
}

// compile-time m where the unrolled polyphase stays within the register budget
template <int M, int R, int MT> struct FastOk { static constexpr bool value = MT * R <= 16; };
#define BTK_MT(MTV) (FastOk<M, R, MTV>::value ? MTV : 0)

template <int M, int R>
static cudaError_t launch_chain_r(const ChainParams& p, int n_work, cudaStream_t st) {
  if (M <= BTK_PP_MAX_M && chain_pp<M, R>(p.m) == 2) {
    constexpr int PP = M <= BTK_PP_MAX_M ? 2 : 1;     // (never instantiated with 2 for the largest transforms)
    if (p.m == 2 && FastOk<M, R, 2>::value) return launch_one<M, R, PP>(btk_chain_kernel<M, R, BTK_MT(2), PP>, p, p.m, n_work, st);
    if (p.m == 4 && FastOk<M, R, 4>::value) return launch_one<M, R, PP>(btk_chain_kernel<M, R, BTK_MT(4), PP>, p, p.m, n_work, st);
    return launch_one<M, R, PP>(btk_chain_kernel<M, R, 0, PP>, p, p.m, n_work, st);
  }
  if (p.m == 2 && FastOk<M, R, 2>::value) return launch_one<M, R, 1>(btk_chain_kernel<M, R, BTK_MT(2), 1>, p, p.m, n_work, st);
  if (p.m == 4 && FastOk<M, R, 4>::value) return launch_one<M, R, 1>(btk_chain_kernel<M, R, BTK_MT(4), 1>, p, p.m, n_work, st);
  return launch_one<M, R, 1>(btk_chain_kernel<M, R, 0, 1>, p, p.m, n_work, st);
}
template <int M, int R>
static cudaError_t launch_analysis_r(const AnalysisParams& p, int n_work, cudaStream_t st) {
  if (p.m == 2 && FastOk<M, R, 2>::value) return launch_one<M, R, 1>(btk_analysis_kernel<M, R, BTK_MT(2)>, p, p.m, n_work, st);
  if (p.m == 4 && FastOk<M, R, 4>::value) return launch_one<M, R, 1>(btk_analysis_kernel<M, R, BTK_MT(4)>, p, p.m, n_work, st);
  return launch_one<M, R, 1>(btk_analysis_kernel<M, R, 0>, p, p.m, n_work, st);
}
template <int M, int R>
static cudaError_t launch_synthesis_r(const SynthesisParams& p, int n_work, cudaStream_t st) {
  if (p.m == 2 && FastOk<M, R, 2>::value) return launch_one<M, R, 1>(btk_synthesis_kernel<M, R, BTK_MT(2)>, p, p.m, n_work, st);
  if (p.m == 4 && FastOk<M, R, 4>::value) return launch_one<M, R, 1>(btk_synthesis_kernel<M, R, BTK_MT(4)>, p, p.m, n_work, st);
  return launch_one<M, R, 1>(btk_synthesis_kernel<M, R, 0>, p, p.m, n_work, st);
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
  int chain_frames_per_iter_m##MM(int R, int m) {                                                                  \
    switch (R) {                                                                                                   \
      case 1: return chain_pp<MM, 1>(m) == 2 ? ChainCfg<MM, 1, 0, 2>::W : ChainCfg<MM, 1, 0, 1>::W;                \
      case 2: return chain_pp<MM, 2>(m) == 2 ? ChainCfg<MM, 2, 0, 2>::W : ChainCfg<MM, 2, 0, 1>::W;                \
      case 4: return chain_pp<MM, 4>(m) == 2 ? ChainCfg<MM, 4, 0, 2>::W : ChainCfg<MM, 4, 0, 1>::W;                \
      case 8: return chain_pp<MM, 8>(m) == 2 ? ChainCfg<MM, 8, 0, 2>::W : ChainCfg<MM, 8, 0, 1>::W;                \
    }                                                                                                              \
    return -1;                                                                                                     \
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
