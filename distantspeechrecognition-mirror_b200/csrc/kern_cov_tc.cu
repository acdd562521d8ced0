// kern_cov_tc.cu -- weighted spatial covariance (SpectralMatrixArray::update, beamformer/beamformer.cc:142-163;
// SubbandBeamformerMVDR.updateSx, lib/subbandBeamforming.py:1170-1175) on the 5th-generation tensor cores.
//
// For one bin s the stage is a dense contraction over frames:  R_s = sum_f wt_f x_f x_f^H  (or x_f x_f^T), x_f in C^C.
// Written over the reals it is ONE symmetric product  P = Z^T Z,  Z = [F][2C] with row f = sqrt(wt_f) * (Re x_0, Im x_0,
// Re x_1, ...) -- exactly the memory order of the snapshots [F][B][C] complex64 -- and
//     R_ab = (P[2a][2b] + P[2a+1][2b+1]) + j (P[2a+1][2b] - P[2a][2b+1])        (x x^H)
//     R_ab = (P[2a][2b] - P[2a+1][2b+1]) + j (P[2a+1][2b] + P[2a][2b+1])        (x x^T)
// so a 64-channel bin is a 128 x 128 x F GEMM whose A and B operands are the SAME shared-memory tile.
//
// tcgen05.mma kind::tf32 keeps 10 mantissa bits of its inputs; to stay inside the 1e-4 parity gate with margin the
// inputs are split  z = hi + lo  (both exactly representable in TF32), and  P = hi hi^T + hi lo^T + lo hi^T  (the
// dropped lo lo^T term is 2^-22 relative).  The last two terms are transposes of each other, so the B operand is the
// stacked tile [hi; lo] (N = 256) and ONE MMA per 8 frames yields  D = [ HH | S ],  HH = hi hi^T,  S = hi lo^T  in 256
// TMEM columns; the epilogue forms  P = HH + S + S^T  (S^T through shared memory).  Against three 128 x 128 MMAs this
// is 2/3 of the tensor time and half of the operand bytes fetched from shared memory, which is what bounds an
// M = N = 128 tile (8 KB of operands per 64 cycles of math = the full 128 B/clk of the shared-memory pipe).
//
// CTA = (bin, frame slice), 256 threads.  Per chunk of 32 frames every thread reads four consecutive frames of four
// real columns (16-byte loads, one 512-byte snapshot row per warp), scales, splits, transposes in registers and writes
// 16-byte words into the two tiles, which are laid out directly in the canonical K-major no-swizzle UMMA layout
// (8 x 16-byte core matrices):
//     tile[k4][row]  (16-byte units),  k4 = frame / 4,  row = rho (hi) or 128 + rho (lo)  ->  LBO (K direction) =
// 4096 B, SBO (8-row groups) = 128 B; A reads rows 0..127, B rows 0..255 of the same slab.  The next chunk's rows are requested before the current one is processed.  Two stages: the MMAs of
// chunk i (issued by one thread, completion signalled through tcgen05.commit on an mbarrier) overlap the staging of
// chunk i+1.  Fewer than 64 channels: several consecutive bins share one tile (see G below).  Epilogue: tcgen05.ld of
// the accumulator (thread = row), S^T through shared memory, pairing of the (Re, Im) rows with one shuffle per column
// pair; one frame slice per tile writes R directly, several slices are merged with fp64 atomics.
#include <stdlib.h>

#include "launch.h"

namespace btk {

#define COV_TC_KC 32                      // frames per stage
#define COV_TC_STAGE_BYTES (256 * COV_TC_KC * 4)   // [KC/4][256 rows] 16-byte words: hi rows 0..127, lo rows 128..255
#define COV_TC_LBO ((256 * 16))                     // bytes between the two 16-byte K groups of one MMA
#define COV_TC_ST_STRIDE 129                       // floats per row of the S^T exchange (conflict-free both ways)
#define COV_TC_SMEM (128 * COV_TC_ST_STRIDE * 4 > 2 * COV_TC_STAGE_BYTES ? 128 * COV_TC_ST_STRIDE * 4 : 2 * COV_TC_STAGE_BYTES)
#define COV_TC_THREADS 256

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

// K-major, no swizzle: start address, leading (K-direction) and stride (8-row group) byte offsets in 16-byte units,
// descriptor version 1 (sm_100), layout type 0.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// D[tmem] (+)= A[smem] * B[smem]^T, M x N x 8 as the instruction descriptor says, TF32 inputs, FP32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// TF32 split with integer arithmetic (cvt.rna.tf32.f32 issues on a slow conversion pipe: 32 of them per thread and
// chunk cost as much as the chunk's MMAs): hi = z rounded to nearest at bit 13 (add half an ulp, clear 13 bits), so
// |z - hi| <= 2^-11 |z|; lo = (z - hi) truncated to TF32 (error <= 2^-11 |lo| <= 2^-22 |z|).
__device__ __forceinline__ float tf32_round(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__global__ void __launch_bounds__(COV_TC_THREADS) btk_covariance_tc_kernel(const cf* __restrict__ snap, const double* __restrict__ wt,
                                                                          double2* __restrict__ Rout, long long F, int B, int C,
                                                                          int conj, int Cp, const CovRec* __restrict__ recs) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (recs) {        // batched launch: blockIdx.z is the recording (its snapshots, recursion weights, frame count, R block)
    const CovRec rc = recs[blockIdx.z];
    snap += rc.snap_off; wt += rc.wt_off; F = rc.F;
    Rout += (long long)blockIdx.z * B * C * C;
  }
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  // Bins per tile: with Cp = C rounded up to a power of two (>= 4) a 128-row tile holds G = 64 / Cp consecutive bins
  // side by side (their snapshot rows are adjacent in memory); the diagonal blocks of P are the G Gram matrices, the
  // cross-bin blocks are computed and dropped -- the tensor pipe has that slack, HBM traffic stays minimal.
  const int G = 64 / Cp;
  const int s = blockIdx.x * G;            // first bin of the tile
  const long long per = (F + gridDim.y - 1) / gridDim.y;
  const long long f_lo = per * blockIdx.y, f_hi = (f_lo + per < F) ? f_lo + per : F;
  const int n_chunks = f_hi > f_lo ? (int)((f_hi - f_lo + COV_TC_KC - 1) / COV_TC_KC) : 0;
  if (n_chunks == 0) return;

  if (tid == 0) {
    mbar_init(smem_u32(&s_bar[0]), 1);
    mbar_init(smem_u32(&s_bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major, N = 256 (>>3 at
  // bit 17), M = 128 (>>4 at bit 24)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t smem_base = smem_u32(smem);

  // Staging map: warp w owns frames 4w .. 4w+3 of the chunk (one 16-byte K group), lane l the real columns 4l .. 4l+3:
  // four coalesced 16-byte loads (one full 512-byte snapshot row per warp) are transposed in registers into four
  // 16-byte stores, one per column.  Column 4l+j is stored as UMMA row rho = (l & 7) + 8 j + 32 (l >> 3): any
  // permutation applied to A and B alike only permutes P, and this one makes the eight lanes of a store phase hit
  // eight different 16-byte bank groups (rows are 16 bytes apart) while keeping the (Re, Im) rows of a channel in the
  // same warp of the epilogue, 8 lanes apart.
  const int lane = tid & 31;
  const bool vec = (C & 1) == 0;           // 8 C bytes per bin row: 16-byte aligned float4 loads need an even C
  const int lg = (4 * lane) / (2 * Cp), lc = (4 * lane) % (2 * Cp);     // this lane's bin of the tile, first real column
  const bool lane_live = s + lg < B;
  const float* zrow = reinterpret_cast<const float*>(snap) + (long long)(s + lg) * C * 2 + lc;
  const long long fstride = (long long)B * C * 2;     // floats between consecutive frames of one bin
  const int rho0 = (lane & 7) + 32 * (lane >> 3);
  auto load4 = [&](long long f) -> float4 {
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f < f_hi && lane_live) {
      const float* p = zrow + f * fstride;
      if (vec) {
        if (lc < 2 * C) x = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        if (lc + 0 < 2 * C) x.x = __ldg(p + 0);
        if (lc + 1 < 2 * C) x.y = __ldg(p + 1);
        if (lc + 2 < 2 * C) x.z = __ldg(p + 2);
        if (lc + 3 < 2 * C) x.w = __ldg(p + 3);
      }
    }
    return x;
  };
  // frame weights travel as raw doubles with the prefetched rows; the square root is taken when the chunk is consumed,
  // so that no instruction waits on a global load inside the prefetch
  float4 xn[4];
  double wn[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const long long f = f_lo + 4 * warp + i;
    xn[i] = load4(f);
    wn[i] = f < f_hi ? __ldg(wt + f) : 0.0;
  }

  for (int ch = 0; ch < n_chunks; ch++) {
    const int st = ch & 1;
    float4 x[4];
    float w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { x[i] = xn[i]; w[i] = sqrtf((float)wn[i]); }
    if (ch + 1 < n_chunks) {               // next chunk's rows are in flight while this one is split and multiplied
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const long long f = f_lo + (long long)(ch + 1) * COV_TC_KC + 4 * warp + i;
        xn[i] = load4(f);
        wn[i] = f < f_hi ? __ldg(wt + f) : 0.0;
      }
    }
    if (ch >= 2) mbar_wait(smem_u32(&s_bar[st]), (uint32_t)(((ch >> 1) - 1) & 1));   // MMAs of chunk ch-2 have read this stage
    float4* hi = reinterpret_cast<float4*>(smem + st * COV_TC_STAGE_BYTES) + warp * 256;
    float4* lo = hi + 128;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float z[4], h4[4], l4[4];
      z[0] = (j == 0 ? x[0].x : j == 1 ? x[0].y : j == 2 ? x[0].z : x[0].w) * w[0];
      z[1] = (j == 0 ? x[1].x : j == 1 ? x[1].y : j == 2 ? x[1].z : x[1].w) * w[1];
      z[2] = (j == 0 ? x[2].x : j == 1 ? x[2].y : j == 2 ? x[2].z : x[2].w) * w[2];
      z[3] = (j == 0 ? x[3].x : j == 1 ? x[3].y : j == 2 ? x[3].z : x[3].w) * w[3];
#pragma unroll
      for (int i = 0; i < 4; i++) { h4[i] = tf32_round(z[i]); l4[i] = tf32_trunc(z[i] - h4[i]); }
      hi[rho0 + 8 * j] = make_float4(h4[0], h4[1], h4[2], h4[3]);
      lo[rho0 + 8 * j] = make_float4(l4[0], l4[1], l4[2], l4[3]);
    }
    // generic-proxy writes -> visible to the tensor core's async proxy, then hand over to the issuing thread
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t slab = smem_base + st * COV_TC_STAGE_BYTES;
#pragma unroll
      for (int k8 = 0; k8 < COV_TC_KC / 8; k8++) {
        // A = rows 0..127 (hi), B = rows 0..255 ([hi; lo]) of the same K slab: D = [hi hi^T | hi lo^T]
        const uint64_t d = umma_desc_kmajor(slab + k8 * 2 * COV_TC_LBO, COV_TC_LBO, 128);
        umma_tf32(tmem, d, d, idesc, (ch | k8) != 0);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar[st]))
                   : "memory");
    }
  }
  // all MMAs complete in issue order: the last chunk's commit covers everything
  {
    const int last = n_chunks - 1;
    mbar_wait(smem_u32(&s_bar[last & 1]), (uint32_t)((last >> 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

#define COV_TC_LD32(u, taddr)                                                                                      \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                    \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                    \
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), \
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),      \
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),     \
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])                   \
      : "r"(taddr)                                                                                                 \
      : "memory");                                                                                                 \
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

  // ---- epilogue.  TMEM lane = UMMA row rho = 32 warp + t (warps 0..3); columns 0..127 hold HH[rho][.], 128..255 S[rho][.].
  // Pass 1: S goes to shared memory (the stage buffers are free: every MMA has completed) so that pass 2 can read S^T.
  float* sT = reinterpret_cast<float*>(smem);
  const int row = 32 * warp + lane;
  if (warp < 4) {
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t u[32];
      COV_TC_LD32(u, tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(128 + c0));
#pragma unroll
      for (int tt = 0; tt < 32; tt++) sT[row * COV_TC_ST_STRIDE + c0 + tt] = __uint_as_float(u[tt]);
    }
  }
  __syncthreads();
  if (warp < 4) {
    // real column of this row = 4 ((t & 7) + 8 warp) + (t >> 3)   (inverse of the staging map)
    const int t = lane;
    const int col = 4 * ((t & 7) + 8 * warp) + (t >> 3);
    const int ga = col / (2 * Cp), a = (col % (2 * Cp)) >> 1, comp = col & 1;   // bin of the tile, channel; even columns
                                             // produce Re R[a][.], odd ones Im R[a][.]; partner row: t ^ 8
    const float sgn = (conj != 0) == (comp == 0) ? 1.f : -1.f;
    double* Rs = reinterpret_cast<double*>(Rout + (long long)(s + ga) * C * C);
    const bool row_live = a < C && s + ga < B;
    const bool single = gridDim.y == 1;      // the only slice of this tile: R is written, not merged
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t u[32], v[32];
      COV_TC_LD32(u, tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0);
      COV_TC_LD32(v, tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(128 + c0));
      float P[32];
#pragma unroll
      for (int tt = 0; tt < 32; tt++)
        P[tt] = __uint_as_float(u[tt]) + (__uint_as_float(v[tt]) + sT[(c0 + tt) * COV_TC_ST_STRIDE + row]);
      // accumulator column n' = c0 + tt holds real column 4 ((tt & 7) + 8 (c0 >> 5)) + (tt >> 3); the (2b, 2b+1) pair
      // of channel b sits at tt and tt + 8 with (tt >> 3) even
#pragma unroll
      for (int tt = 0; tt < 32; tt++) {
        if ((tt >> 3) & 1) continue;
        const float v0 = P[tt], v1 = P[tt + 8];
        const float pv1 = __shfl_xor_sync(0xffffffffu, v1, 8);
        const int colb = 4 * ((tt & 7) + 8 * (c0 >> 5)) + (tt >> 3);      // even real column of the pair
        const int gb = colb / (2 * Cp), b = (colb % (2 * Cp)) >> 1;
        if (row_live && gb == ga && b < C) {
          double* dst = Rs + 2 * ((long long)a * C + b) + comp;
          const double val = (double)(v0 + sgn * pv1);
          if (single) *dst = val; else atomicAdd(dst, val);   // (the caller zeroes R: a single slice simply writes it)
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

cudaError_t launch_covariance_tc(const cf* snap, const double* wt, double2* Rout, long long F, int B, int C, int conj,
                                 cudaStream_t st) {
  if (C > 64 || C < 1) return cudaErrorInvalidValue;
  if (F == 0) return cudaSuccess;
  int Cp = 4;
  while (Cp < C) Cp *= 2;
  const int tiles = (B + 64 / Cp - 1) / (64 / Cp);
  // frame slices: at most 2048 frames each (FP32 accumulation span in TMEM).  With at least one tile per SM a single
  // slice per tile is best (R is then written without atomics); fewer tiles are cut into slices of at least 64 frames
  // until there are about 1.5 CTAs per SM (measured: more slices only add atomics).
  int split = (int)((F + 2047) / 2048);
  const int want = tiles >= 148 ? 1 : (3 * 148 / 2 + tiles - 1) / tiles;   // measured: C = 64 best unsplit, C = 16 best at 4 slices
  if (split < want) split = want;
  const int cap = (int)((F + 63) / 64);
  if (split > cap) split = cap;
  if (split < 1) split = 1;
  // (no upper cap other than the grid limit: F > 131072 frames must still be cut into slices of at most 2048 frames, the
  // documented FP32 accumulation span; the slices are merged with fp64 atomics)
  if (split > 65535) split = 65535;
  if (const char* e = getenv("BTK_COV_SPLIT")) { const int v = atoi(e); if (v >= 1 && v <= 64) split = v; }   // tuning knob (A/B runs)
  const int smem = COV_TC_SMEM;
  cudaError_t e = cudaFuncSetAttribute(btk_covariance_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  btk_covariance_tc_kernel<<<dim3(tiles, split), COV_TC_THREADS, smem, st>>>(snap, wt, Rout, F, B, C, conj, Cp, nullptr);
  return cudaGetLastError();
}

// n recordings in one launch (grid.z): recs[i] = (snapshot offset, weight offset, frames), R block i = Rout + i B C C.
// One frame slice per (tile, recording) as soon as the grid fills the SMs (R is then written without atomics and needs no
// zeroing); Fmax = the largest frame count of the batch.
cudaError_t launch_covariance_tc_batch(const cf* snap, const double* wt, double2* Rout, const CovRec* recs, int n, long long Fmax,
                                       int B, int C, int conj, cudaStream_t st) {
  if (C > 64 || C < 1 || n < 1 || n > 65535) return cudaErrorInvalidValue;
  if (Fmax == 0) return cudaSuccess;
  int Cp = 4;
  while (Cp < C) Cp *= 2;
  const int tiles = (B + 64 / Cp - 1) / (64 / Cp);
  int split = (int)((Fmax + 2047) / 2048);
  const long long ctas = (long long)tiles * n;
  const int want = ctas >= 148 ? 1 : (int)((3 * 148 / 2 + ctas - 1) / ctas);
  if (split < want) split = want;
  const int cap = (int)((Fmax + 63) / 64);
  if (split > cap) split = cap;
  if (split < 1) split = 1;
  if (split > 64) split = 64;
  cudaError_t e = cudaFuncSetAttribute(btk_covariance_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, COV_TC_SMEM);
  if (e != cudaSuccess) return e;
  btk_covariance_tc_kernel<<<dim3(tiles, split, n), COV_TC_THREADS, COV_TC_SMEM, st>>>(snap, wt, Rout, 0, B, C, conj, Cp, recs);
  return cudaGetLastError();
}

}  // namespace btk
