// kern_misc.cu -- per-M dispatch glue + the subband-domain kernels that are not filter banks:
// weight apply on stored snapshots, weighted spatial covariance accumulation, per-bin MVDR solve.
#include <stdlib.h>

#include "launch.h"
#include "snap_tile.cuh"

namespace btk {

#define BTK_DECL_M(MM)                                                                                \
  cudaError_t launch_chain_m##MM(int R, const ChainParams& p, int n_work, cudaStream_t st);           \
  cudaError_t launch_analysis_m##MM(int R, const AnalysisParams& p, int n_work, cudaStream_t st);     \
  cudaError_t launch_synthesis_m##MM(int R, const SynthesisParams& p, int n_work, cudaStream_t st);   \
  int fb_smem_bytes_m##MM(int R, int m);                                                              \
  int chain_frames_per_iter_m##MM(int R, int m);
BTK_DECL_M(64) BTK_DECL_M(128) BTK_DECL_M(256) BTK_DECL_M(512) BTK_DECL_M(1024)
#undef BTK_DECL_M

// warp-specialised fused chain (kern_ws.cuh), one translation unit per transform size
#define BTK_DECL_WS(MM)                                                                               \
  cudaError_t launch_chain_ws_m##MM(int R, const ChainParams& p, int n_work, cudaStream_t st);        \
  int chain_ws_frames_per_iter_m##MM(int R, int m);                                                   \
  bool chain_ws_cluster_ok_m##MM(int R, int m, int S);
BTK_DECL_WS(64) BTK_DECL_WS(128) BTK_DECL_WS(256) BTK_DECL_WS(512) BTK_DECL_WS(1024)
#undef BTK_DECL_WS

bool fb_supported(int M, int R) {
  if (R != 1 && R != 2 && R != 4 && R != 8) return false;
  switch (M) {
    case 64: case 128: return R <= 8;       // first radix 8
    case 256: case 512: case 1024: return R <= 8;
  }
  return false;
}

#define BTK_DISPATCH(fn, ...)                               \
  switch (M) {                                              \
    case 64: return fn##_m64(__VA_ARGS__);                  \
    case 128: return fn##_m128(__VA_ARGS__);                \
    case 256: return fn##_m256(__VA_ARGS__);                \
    case 512: return fn##_m512(__VA_ARGS__);                \
    case 1024: return fn##_m1024(__VA_ARGS__);              \
  }

cudaError_t launch_chain(int M, int R, const ChainParams& p, int n_work, cudaStream_t st) {
  if (!fb_supported(M, R)) return cudaErrorInvalidValue;
  BTK_DISPATCH(launch_chain, R, p, n_work, st)
  return cudaErrorInvalidValue;
}
cudaError_t launch_analysis(int M, int R, const AnalysisParams& p, int n_work, cudaStream_t st) {
  if (!fb_supported(M, R)) return cudaErrorInvalidValue;
  BTK_DISPATCH(launch_analysis, R, p, n_work, st)
  return cudaErrorInvalidValue;
}
cudaError_t launch_synthesis(int M, int R, const SynthesisParams& p, int n_work, cudaStream_t st) {
  if (!fb_supported(M, R)) return cudaErrorInvalidValue;
  BTK_DISPATCH(launch_synthesis, R, p, n_work, st)
  return cudaErrorInvalidValue;
}
int fb_smem_bytes(int M, int R, int m) {
  if (!fb_supported(M, R)) return -1;
  BTK_DISPATCH(fb_smem_bytes, R, m)
  return -1;
}
cudaError_t launch_chain_ws(int M, int R, const ChainParams& p, int n_work, cudaStream_t st) {
  if (!fb_supported(M, R)) return cudaErrorInvalidValue;
  BTK_DISPATCH(launch_chain_ws, R, p, n_work, st)
  return cudaErrorInvalidValue;
}
int chain_ws_frames_per_iter(int M, int R, int m) {
  if (!fb_supported(M, R)) return -1;
  BTK_DISPATCH(chain_ws_frames_per_iter, R, m)
  return -1;
}
bool chain_ws_cluster_ok(int M, int R, int m, int S) {
  if (!fb_supported(M, R)) return false;
  BTK_DISPATCH(chain_ws_cluster_ok, R, m, S)
  return false;
}
int fb_frames_per_iter(int M, int R) { return (M >= 1024 || (M == 512 && R == 1)) ? 8 : 16; }  // ChainCfg::W = 2 * NW (staged kernels)
int chain_frames_per_iter(int M, int R, int m) {                   // the fused chain may window two pairs per warp
  if (!fb_supported(M, R)) return -1;
  BTK_DISPATCH(chain_frames_per_iter, R, m)
  return -1;
}

// ---------------------------------------------------------------------------------------------
// Weight apply on stored snapshots (SubbandDS::next / SubbandMVDR::next zdotc loop, beamformer.cc:1181-1194, 2616-2630):
// a lane group per (frame, bin) item, channels across the lanes (snap_tile.cuh); HBM-bound: 8 C F B + 8 F B bytes.
// ---------------------------------------------------------------------------------------------
template <int GS>
__global__ void __launch_bounds__(SNAP_THREADS) btk_beamform_kernel(const cf* __restrict__ snap, const cf* __restrict__ w,
                                                                   cf* __restrict__ Y, long long FB, int B, int C) {
  const int lg = threadIdx.x % GS;
  const long long g0 = ((long long)blockIdx.x * SNAP_THREADS + threadIdx.x) / GS;
  const long long gstride = (long long)gridDim.x * (SNAP_THREADS / GS);
  const long long n_it = (FB + gstride - 1) / gstride;        // the same trip count for every lane of a warp
  constexpr int U = 4;                                        // items in flight per group (independent loads)
  for (long long it = 0; it < n_it; it += U) {
    float re[U], im[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long idx = g0 + (it + u) * gstride;
      re[u] = 0.f; im[u] = 0.f;
      if (it + u < n_it && idx < FB) {
        const int s = (int)(idx % B);
        const cf* x = snap + idx * C;
        const cf* ws = w + (long long)s * C;
        for (int c = lg; c < C; c += GS) {
          const cf a = __ldg(ws + c), b = x[c];     // conj(a) * b
          re[u] = fmaf(a.x, b.x, re[u]); re[u] = fmaf(a.y, b.y, re[u]);
          im[u] = fmaf(a.x, b.y, im[u]); im[u] = fmaf(-a.y, b.x, im[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long idx = g0 + (it + u) * gstride;
      const float r = group_sum<GS>(re[u]), i = group_sum<GS>(im[u]);
      if (it + u < n_it && idx < FB && lg == 0) Y[idx] = mk(r, i);
    }
  }
}

cudaError_t launch_beamform(const cf* snap, const cf* w, cf* Y, long long F, int B, int C, cudaStream_t st) {
  const long long FB = F * B;
  if (FB == 0) return cudaSuccess;
  const int grid = snap_grid(FB, C);
  switch (snap_group_size(C)) {
    case 1: btk_beamform_kernel<1><<<grid, SNAP_THREADS, 0, st>>>(snap, w, Y, FB, B, C); break;
    case 2: btk_beamform_kernel<2><<<grid, SNAP_THREADS, 0, st>>>(snap, w, Y, FB, B, C); break;
    case 4: btk_beamform_kernel<4><<<grid, SNAP_THREADS, 0, st>>>(snap, w, Y, FB, B, C); break;
    case 8: btk_beamform_kernel<8><<<grid, SNAP_THREADS, 0, st>>>(snap, w, Y, FB, B, C); break;
    case 16: btk_beamform_kernel<16><<<grid, SNAP_THREADS, 0, st>>>(snap, w, Y, FB, B, C); break;
    default: btk_beamform_kernel<32><<<grid, SNAP_THREADS, 0, st>>>(snap, w, Y, FB, B, C); break;
  }
  return cudaGetLastError();
}

__device__ __forceinline__ double shfl_d(double v, int src);

// ---------------------------------------------------------------------------------------------
// Pieces of the device-resident MVDR adaptation (btkb200_mvdr_chain_batch): diagonal loading of the estimated
// matrices and the chain weight table, both without a trip to the host.
// ---------------------------------------------------------------------------------------------
// R[s][c][c] += (double)(float)load_abs + load_rel * trace(R[s]) / C   (setAllLevelsOfDiagonalLoading keeps the weight as a
// float, beamformer.cc:2342, 2562-2565; the relative term is the "1e-2 trace(R)/C" loading of SURVEY 8d).  One warp per bin.
__global__ void btk_diag_load_kernel(double2* __restrict__ Rn, int B, int C, float load_abs, double load_rel) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s >= B) return;
  double2* Rs = Rn + (long long)s * C * C;
  double tr = 0.0;
  for (int c = lane; c < C; c += 32) tr += Rs[c * C + c].x;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tr += shfl_d(tr, lane ^ o);
  const double add = (double)load_abs + load_rel * tr / (double)C;
  for (int c = lane; c < C; c += 32) Rs[c * C + c].x += add;
}

// Hermitian-extended conjugate weight table of the fused chain from double weights on the device (the device twin of
// host_tables.h::build_chain_weight_table; binmap[slot] = bin of table slot ((r/2) L + gl) 2 + (r & 1)).
__global__ void btk_weight_table_kernel(const double2* __restrict__ w, const int* __restrict__ binmap, cf* __restrict__ gam, int M, int C,
                                        int Cpad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cpad * M) return;
  w += (long long)blockIdx.y * (M / 2 + 1) * C;          // grid.y = recording of a batch
  gam += (long long)blockIdx.y * Cpad * M;
  const int c = idx / M, slot = idx - c * M;
  cf v = mk(0.f, 0.f);
  if (c < C) {
    const int k = binmap[slot];
    if (k == 0 || k == M / 2) v = mk((float)w[(long long)k * C + c].x, 0.f);
    else if (k < M / 2) { const double2 z = w[(long long)k * C + c]; v = mk((float)z.x, (float)-z.y); }
    else { const double2 z = w[(long long)(M - k) * C + c]; v = mk((float)z.x, (float)z.y); }
  }
  gam[idx] = v;
}

cudaError_t launch_diag_load(double2* Rn, int B, int C, float load_abs, double load_rel, cudaStream_t st) {
  btk_diag_load_kernel<<<(B + 3) / 4, 128, 0, st>>>(Rn, B, C, load_abs, load_rel);
  return cudaGetLastError();
}
cudaError_t launch_weight_table(const double2* w, const int* binmap, cf* gam, int M, int C, int Cpad, cudaStream_t st, int n) {
  btk_weight_table_kernel<<<dim3((Cpad * M + 255) / 256, n), 256, 0, st>>>(w, binmap, gam, M, C, Cpad);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Ingest: raw interleaved PCM -> float32 (same [T][C] order).  HBM-bound element-wise converts; four samples per
// thread per step (8- or 12-byte loads, one 16-byte store).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) btk_ingest_s16_kernel(const short* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const short4 v = reinterpret_cast<const short4*>(src)[i];
    reinterpret_cast<float4*>(dst)[i] = make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[(n4 << 2) + threadIdx.x] = (float)src[(n4 << 2) + threadIdx.x];
}

__device__ __forceinline__ float s24be(unsigned b0, unsigned b1, unsigned b2) {   // b0 = most significant byte
  return (float)((int)((b0 << 24) | (b1 << 16) | (b2 << 8)) >> 8);
}

__global__ void __launch_bounds__(256) btk_ingest_s24be_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst,
                                                              long long n) {
  const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const unsigned* p = reinterpret_cast<const unsigned*>(src) + 3 * i;      // 12 bytes = 4 samples
    const unsigned w0 = p[0], w1 = p[1], w2 = p[2];                          // little-endian words of the byte stream
    float4 o;
    o.x = s24be(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
    o.y = s24be(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
    o.z = s24be((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
    o.w = s24be((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
    reinterpret_cast<float4*>(dst)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const unsigned char* q = src + 3 * ((n4 << 2) + threadIdx.x);
    dst[(n4 << 2) + threadIdx.x] = s24be(q[0], q[1], q[2]);
  }
}

cudaError_t launch_ingest(int fmt, const void* src, float* dst, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  long long blocks = ((n >> 2) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (fmt == 1) btk_ingest_s16_kernel<<<(int)blocks, 256, 0, st>>>((const short*)src, dst, n);
  else if (fmt == 2) btk_ingest_s24be_kernel<<<(int)blocks, 256, 0, st>>>((const unsigned char*)src, dst, n);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Weighted Gram matrices R[s] = sum_f wt[f] x_f x_f^H (conj) or x_f x_f^T, register-tiled outer products.
// grid = (B bins, SPLIT frame slices).  A CTA stages 32 frames of its bin in shared memory twice -- x and
// wt*conj(x) -- and every thread owns one TILE x TILE block of the upper triangle (both flavours are symmetric
// up to conjugation): per frame 2*TILE complex loads feed TILE^2 packed complex multiply-adds.  fp32 partial sums
// over <= 32 frames are promoted to fp64 accumulators; slices are merged with fp64 atomics into Rout (zeroed by the
// caller), the lower triangle is written as the mirror.
// ---------------------------------------------------------------------------------------------
#define BTK_COV_FB 32
template <int TILE>
__global__ void __launch_bounds__(256) btk_covariance_kernel(const cf* __restrict__ snap, const double* __restrict__ wt,
                                                            double2* __restrict__ Rout, long long F, int B, int C,
                                                            int conj) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Cp = (C + TILE - 1) / TILE * TILE, nt = Cp / TILE, ntile = nt * (nt + 1) / 2;
  cf* sx = reinterpret_cast<cf*>(smem_raw);              // [FB][Cp]  x
  cf* sy = sx + BTK_COV_FB * Cp;                          // [FB][Cp]  wt * (conj ? conj(x) : x)
  const int s = blockIdx.x;
  const long long per = (F + gridDim.y - 1) / gridDim.y;
  const long long f_lo = per * blockIdx.y, f_hi = (f_lo + per < F) ? f_lo + per : F;
  // upper-triangular tile (ti <= tj) of this thread
  int ti = 0, tj = 0;
  const bool active = (int)threadIdx.x < ntile;
  {
    int kk = threadIdx.x;
    for (ti = 0; ti < nt; ti++) {
      const int len = nt - ti;
      if (kk < len) { tj = ti + kk; break; }
      kk -= len;
    }
    if (!active) { ti = 0; tj = 0; }
  }
  double2 acc[TILE][TILE];
#pragma unroll
  for (int p = 0; p < TILE; p++)
#pragma unroll
    for (int q = 0; q < TILE; q++) acc[p][q] = make_double2(0.0, 0.0);
  for (long long f0 = f_lo; f0 < f_hi; f0 += BTK_COV_FB) {
    const int nf = (int)((f_hi - f0 < BTK_COV_FB) ? f_hi - f0 : BTK_COV_FB);
    __syncthreads();
    for (int i = threadIdx.x; i < nf * Cp; i += blockDim.x) {
      const int ff = i / Cp, c = i % Cp;
      cf v = mk(0.f, 0.f);
      if (c < C) v = snap[((f0 + ff) * B + s) * C + c];
      const float w = (float)wt[f0 + ff];
      sx[i] = v;
      sy[i] = mk(w * v.x, conj ? -w * v.y : w * v.y);
    }
    __syncthreads();
    if (active) {
      cf part[TILE][TILE];
#pragma unroll
      for (int p = 0; p < TILE; p++)
#pragma unroll
        for (int q = 0; q < TILE; q++) part[p][q] = mk(0.f, 0.f);
      const cf* pa = sx + ti * TILE;
      const cf* pb = sy + tj * TILE;
      for (int ff = 0; ff < nf; ff++) {
        cf a[TILE], b[TILE];
#pragma unroll
        for (int p = 0; p < TILE; p += 2) {
          const float4 va = *reinterpret_cast<const float4*>(pa + ff * Cp + p);
          const float4 vb = *reinterpret_cast<const float4*>(pb + ff * Cp + p);
          a[p] = mk(va.x, va.y); a[p + 1] = mk(va.z, va.w);
          b[p] = mk(vb.x, vb.y); b[p + 1] = mk(vb.z, vb.w);
        }
#pragma unroll
        for (int p = 0; p < TILE; p++)
#pragma unroll
          for (int q = 0; q < TILE; q++) cfma(part[p][q], a[p], b[q]);
      }
#pragma unroll
      for (int p = 0; p < TILE; p++)
#pragma unroll
        for (int q = 0; q < TILE; q++) { acc[p][q].x += (double)part[p][q].x; acc[p][q].y += (double)part[p][q].y; }
    }
  }
  if (active) {
    double2* Rs = Rout + (long long)s * C * C;
#pragma unroll
    for (int p = 0; p < TILE; p++)
#pragma unroll
      for (int q = 0; q < TILE; q++) {
        const int i = ti * TILE + p, j = tj * TILE + q;
        if (i >= C || j >= C) continue;
        atomicAdd(&Rs[i * C + j].x, acc[p][q].x);
        atomicAdd(&Rs[i * C + j].y, acc[p][q].y);
        if (ti != tj) {      // mirror: R_ji = conj(R_ij) (Hermitian flavour) or R_ij (x x^T flavour)
          atomicAdd(&Rs[j * C + i].x, acc[p][q].x);
          atomicAdd(&Rs[j * C + i].y, conj ? -acc[p][q].y : acc[p][q].y);
        }
      }
  }
}

cudaError_t launch_covariance(const cf* snap, const double* wt, double2* Rout, long long F, int B, int C, int conj,
                              cudaStream_t st) {
  if (C > 64) return cudaErrorInvalidValue;
  if (F == 0) return cudaSuccess;
  // The per-bin contraction (2C x 2C x F over the reals) runs on the tensor cores (kern_cov_tc.cu; bins are packed
  // side by side into the 128 x 128 tcgen05 tile when C < 64).  BTK_COV_SIMT=1 selects the register-tiled SIMT kernel
  // below instead (A/B measurements, tests of both paths).
  static const bool force_simt = getenv("BTK_COV_SIMT") && getenv("BTK_COV_SIMT")[0] == '1';
  if (!force_simt) return launch_covariance_tc(snap, wt, Rout, F, B, C, conj, st);
  int split = (int)((F + 511) / 512);
  if (split < 1) split = 1;
  if (split > 64) split = 64;
  if (C >= 32) {
    const int Cp = (C + 3) / 4 * 4;
    const size_t smem = (size_t)2 * BTK_COV_FB * Cp * sizeof(cf);
    btk_covariance_kernel<4><<<dim3(B, split), 160, smem, st>>>(snap, wt, Rout, F, B, C, conj);
  } else {
    const int Cp = (C + 1) / 2 * 2;
    const size_t smem = (size_t)2 * BTK_COV_FB * Cp * sizeof(cf);
    btk_covariance_kernel<2><<<dim3(B, split), 160, smem, st>>>(snap, wt, Rout, F, B, C, conj);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// MVDR weights, one CTA per bin (beamformer.cc:2392-2446):  t = (R^H)^{-1} d ; lam = t^H d ; w = t/(lam C).
// Pivoted Gaussian elimination on the augmented system [R^H | d] in complex double in shared memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 zmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 zdiv(double2 a, double2 b) {
  const double d = b.x * b.x + b.y * b.y;
  return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

__device__ __forceinline__ double shfl_d(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src); hi = __shfl_sync(0xffffffffu, hi, src);
  return __hiloint2double(hi, lo);
}

__global__ void __launch_bounds__(256) btk_mvdr_solve_kernel(const double2* __restrict__ Rn, const double2* __restrict__ dvec,
                                                            double2* __restrict__ w, int* __restrict__ fallback, int C,
                                                            double dThreshold, int B) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* A = reinterpret_cast<double2*>(smem_raw);       // [C][C+1]
  __shared__ double s_best[8];
  __shared__ int s_bidx[8];
  // block = (recording, bin): the manifold is shared by the recordings of a batch, R / w / fallback are per recording
  const int sb = blockIdx.x, s = sb % B, ld = C + 1, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const double2* d = dvec + (long long)s * C;
  double2* ws = w + (long long)sb * C;
  fallback += sb - s;                          // this recording's flags
  if (s == 0) {   // w[0] = (1, ..., 1), beamformer.cc:2410-2415
    for (int c = tid; c < C; c += nt) ws[c] = make_double2(1.0, 0.0);
    if (tid == 0) fallback[0] = 0;
    return;
  }
  const double2* Rs = Rn + (long long)sb * C * C;
  for (int i = tid; i < C * C; i += nt) {
    const int r = i / C, c = i % C;
    const double2 v = Rs[c * C + r];                        // (R^H)[r][c] = conj(R[c][r])
    A[r * ld + c] = make_double2(v.x, -v.y);
  }
  for (int r = tid; r < C; r += nt) A[r * ld + C] = d[r];
  bool bad = false;
  __syncthreads();
  for (int k = 0; k < C; k++) {
    // partial pivoting: arg max |A[r][k]|^2 over r >= k, one row per thread (C <= 128)
    double best = -1.0; int bidx = k;
    if (tid >= k && tid < C) { const double2 v = A[tid * ld + k]; best = v.x * v.x + v.y * v.y; bidx = tid; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = shfl_d(best, lane ^ o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (lane == 0) { s_best[warp] = best; s_bidx[warp] = bidx; }
    __syncthreads();
    // every thread folds the per-warp candidates itself (broadcast loads, same result everywhere): no serial section
    // and no second barrier; the slots are rewritten two barriers later
    double bb = s_best[0]; int piv = s_bidx[0];
    for (int q = 1; q < (nt >> 5); q++) { const double b2 = s_best[q]; const int i2 = s_bidx[q]; if (b2 > bb || (b2 == bb && i2 < piv)) { bb = b2; piv = i2; } }
    if (!(bb > dThreshold * dThreshold) || !isfinite(bb)) { bad = true; break; }
    if (piv != k) {
      for (int c = k + tid; c <= C; c += nt) {
        const double2 t = A[k * ld + c]; A[k * ld + c] = A[piv * ld + c]; A[piv * ld + c] = t;
      }
    }
    __syncthreads();
    // rank-1 update of rows k+1.., columns k+1..C (incl. the rhs): a warp per row, lanes across the columns
    // (consecutive 16-byte words: conflict-free; the pivot-row elements of a lane's columns stay in registers, the row
    // factor f_r = A[r][k] / pivot is a broadcast load + one multiply per lane -- no separate scaling pass, no
    // index division)
    {
      // 1 / pivot = conj(pivot) / |pivot|^2 with one reciprocal (a full complex division costs two fp64 divisions)
      const double2 pv = A[k * ld + k];
      const double rd = __drcp_rn(pv.x * pv.x + pv.y * pv.y);
      const double2 inv = make_double2(pv.x * rd, -pv.y * rd);
      double2 pk[3];                                  // columns k+1+lane, +32, +64  (C <= 64 -> at most 65 columns)
#pragma unroll
      for (int q = 0; q < 3; q++) { const int c = k + 1 + lane + 32 * q; pk[q] = c <= C ? A[k * ld + c] : make_double2(0.0, 0.0); }
      for (int r = k + 1 + warp; r < C; r += (nt >> 5)) {
        const double2 f = zmul(A[r * ld + k], inv);
#pragma unroll
        for (int q = 0; q < 3; q++) {
          const int c = k + 1 + lane + 32 * q;
          if (c <= C) {
            const double2 t = zmul(f, pk[q]);
            double2 a = A[r * ld + c];
            a.x -= t.x; a.y -= t.y;
            A[r * ld + c] = a;
          }
        }
      }
    }
    __syncthreads();
  }
  // ---- the reference's rejection rule (pseudoinverse, beamformer.cc:275-283): ANY singular value below dThreshold makes the
  // inverse "fail" and the caller falls back to the identity (:2425-2427).  A pivot is not a singular value, so
  // the smallest singular value is estimated from the triangular factor: A = P^T L U with |L_ij| <= 1, hence
  // sigma_min(A) ~ sigma_min(U) up to the (modest) conditioning of L, and sigma_min(U) = 1 / sqrt(lambda_max((U^H U)^-1)) comes
  // from four steps of inverse iteration -- two triangular solves each, inside warp 0, rows of U read along the lanes.
  // Where sigma_min sits within a factor of a few of the threshold the reference's own answer is rounding noise of its
  // single-precision SVD; away from it the two rules agree (tests/test_mvdr_fallback.py, against the compiled reference).
  if (!bad && warp == 0 && dThreshold > 0.0) {
    double2 x[4];
    double nrm2 = 0.0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int i = lane + 32 * q;
      x[q] = i < C ? make_double2((i & 1) ? -1.0 - (double)i / C : 1.0 + (double)i / C, 0.0) : make_double2(0.0, 0.0);
      nrm2 += x[q].x * x[q].x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm2 += shfl_d(nrm2, lane ^ o);
    double lam = 0.0;
    for (int it = 0; it < 4; it++) {
      const double sc = rsqrt(nrm2);
#pragma unroll
      for (int q = 0; q < 4; q++) { x[q].x *= sc; x[q].y *= sc; }
      // forward substitution with U^H (lower triangular, (U^H)_ir = conj(U_ri)): y_r = x_r / conj(U_rr); x_i -= conj(U_ri) y_r
      for (int r = 0; r < C; r++) {
        double2 yr = make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < 4; q++) if ((r >> 5) == q && lane == (r & 31)) { const double2 u = A[r * ld + r]; yr = zdiv(x[q], make_double2(u.x, -u.y)); }
        yr.x = shfl_d(yr.x, r & 31); yr.y = shfl_d(yr.y, r & 31);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int i = lane + 32 * q;
          if (i > r && i < C) { const double2 u = A[r * ld + i]; const double2 t = zmul(make_double2(u.x, -u.y), yr); x[q].x -= t.x; x[q].y -= t.y; }
          if (i == r) x[q] = yr;
        }
      }
      // back substitution with U
      for (int r = C - 1; r >= 0; r--) {
        double2 zr = make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < 4; q++) if ((r >> 5) == q && lane == (r & 31)) zr = zdiv(x[q], A[r * ld + r]);
        zr.x = shfl_d(zr.x, r & 31); zr.y = shfl_d(zr.y, r & 31);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int i = lane + 32 * q;
          if (i < r) { const double2 t = zmul(A[i * ld + r], zr); x[q].x -= t.x; x[q].y -= t.y; }
          if (i == r) x[q] = zr;
        }
      }
      nrm2 = 0.0;
#pragma unroll
      for (int q = 0; q < 4; q++) nrm2 += x[q].x * x[q].x + x[q].y * x[q].y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nrm2 += shfl_d(nrm2, lane ^ o);
      lam = sqrt(nrm2);                               // ||(U^H U)^-1 x||, x of unit length -> lambda_max estimate (from below)
      if (!isfinite(lam)) break;
    }
    // sigma_min(U) ~ 1 / sqrt(lam)
    if (!isfinite(lam) || !(lam * dThreshold * dThreshold < 1.0)) s_bidx[0] = -1; else s_bidx[0] = 0;
  }
  if (!bad && dThreshold > 0.0) {
    __syncthreads();
    if (s_bidx[0] < 0) bad = true;
  }
  if (bad) {
    // identity fallback (beamformer.cc:2425-2427): t = d
    __syncthreads();
    if (tid == 0) {
      double lr = 0.0;
      for (int c = 0; c < C; c++) lr += d[c].x * d[c].x + d[c].y * d[c].y;
      for (int c = 0; c < C; c++) ws[c] = make_double2(d[c].x / (lr * C), d[c].y / (lr * C));
      fallback[s] = 1;
    }
    return;
  }
  if (warp == 0) {
    // back substitution, column oriented, inside one warp: lane owns rows lane, lane+32, ... (right-hand sides in
    // registers); t_r is broadcast with shuffles, no block barrier
    double2 rhs[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { const int r = lane + 32 * q; rhs[q] = r < C ? A[r * ld + C] : make_double2(0.0, 0.0); }
    for (int r = C - 1; r >= 0; r--) {
      double2 tr = make_double2(0.0, 0.0);
#pragma unroll
      for (int q = 0; q < 4; q++) if ((r >> 5) == q && lane == (r & 31)) tr = zdiv(rhs[q], A[r * ld + r]);
      tr.x = shfl_d(tr.x, r & 31); tr.y = shfl_d(tr.y, r & 31);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int i = lane + 32 * q;
        if (i < r) { const double2 t = zmul(A[i * ld + r], tr); rhs[q].x -= t.x; rhs[q].y -= t.y; }
        if (i == r) rhs[q] = tr;                      // t_r overwrites the rhs
      }
    }
    double2 lam = make_double2(0.0, 0.0);                  // lam = t^H d  (gsl_blas_zdotc(tmpH, d))
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int c = lane + 32 * q;
      if (c < C) { lam.x += rhs[q].x * d[c].x + rhs[q].y * d[c].y; lam.y += rhs[q].x * d[c].y - rhs[q].y * d[c].x; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lam.x += shfl_d(lam.x, lane ^ o); lam.y += shfl_d(lam.y, lane ^ o); }
    const double2 nrm = make_double2(lam.x * C, lam.y * C);
#pragma unroll
    for (int q = 0; q < 4; q++) { const int c = lane + 32 * q; if (c < C) ws[c] = zdiv(rhs[q], nrm); }
    if (lane == 0) fallback[s] = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// MVDR weights for HERMITIAN matrices (the adaptive batch: x x^H covariances + diagonal loading), one CTA of 128 threads per
// (recording, bin), C <= 64.  Same result as btk_mvdr_solve_kernel (beamformer.cc:2392-2446:  t = (R^H)^-1 d, lam = t^H d,
// w = t / (lam C), identity fallback when pseudoinverse() would drop a singular value, :275-283 / :2425-2427), different
// route: the general kernel is a pivoted LU with three CTA barriers per column, four steps of inverse iteration in one warp
// and a complex division inside every step of its triangular solves -- 1 us per bin at 64 channels, 263 ms for the 1 024
// utterances of BASELINE configs[3], eight times their chain.  A Hermitian positive definite matrix needs none of that:
//   * R = L D L^H without pivoting on the packed lower triangle (35 KB at 64 channels: five CTAs per SM), in PANELS of
//     eight columns: the 8 x 8 diagonal block inside one warp, the rows below it one thread each with their panel entries
//     in registers, then ONE rank-8 update of the trailing matrix (one read-modify-write of an entry per eight
//     multiply-adds, the multipliers conj(c_jp) / D_p broadcast from a small table) -- three CTA barriers per panel;
//   * the right-hand sides ride along as two extra ROWS of the matrix: row C = d^H gives the forward substitution for t,
//     row C + 1 = x0^H the one for the singular-value probe, both finished when the factorisation is;
//   * one back substitution in warp 0 for both, D real: reciprocals are taken once per column, no division in the chain;
//   * sigma_min(R) <= |x| / |R^-1 x| for x = x0 and x = d (one step of inverse iteration each): below dThreshold the bin
//     is rejected, as is a pivot that is not positive -- the semidefinite covariance of fewer frames than channels ends
//     there, which is the bin the reference's SVD rejects too.
// ---------------------------------------------------------------------------------------------
#define BTK_CHOL_PANEL 8
__device__ __forceinline__ int chol_row_off(int i, int C) { return i <= C ? (i * (i + 1)) / 2 : (C * (C + 1)) / 2 + (i - C) * C; }

__global__ void __launch_bounds__(128) btk_mvdr_chol_kernel(const double2* __restrict__ Rn, const double2* __restrict__ dvec,
                                                           double2* __restrict__ w, int* __restrict__ fallback, int C,
                                                           double dThreshold, int B) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NP = BTK_CHOL_PANEL;
  const int NR = C + 2;                                          // matrix rows + the two right-hand-side rows
  double2* A = reinterpret_cast<double2*>(smem_raw);             // packed rows: (i, j), j <= min(i, C - 1)
  double2* Wm = A + chol_row_off(NR, C);                         // [NR][NP] multipliers conj(c_jp) / D_p of the current panel
  double2* Bm = Wm + NR * NP;                                    // [NP][NP] conj(L_pq) of the panel's diagonal block
  double* Dinv = reinterpret_cast<double*>(Bm + NP * NP);        // [C]
  __shared__ int s_bad;
  const int sb = blockIdx.x, s = sb % B, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const double2* d = dvec + (long long)s * C;
  double2* ws = w + (long long)sb * C;
  fallback += sb - s;
  if (s == 0) {   // w[0] = (1, ..., 1), beamformer.cc:2410-2415
    for (int c = tid; c < C; c += nt) ws[c] = make_double2(1.0, 0.0);
    if (tid == 0) fallback[0] = 0;
    return;
  }
  const double2* Rs = Rn + (long long)sb * C * C;
  // lower triangle of R (= R^H for a Hermitian R): coalesced walk over the full rows, eight loads in flight per thread
  for (int e0 = 0; e0 < C * C; e0 += 8 * nt) {
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { const int e = e0 + u * nt + tid; v[u] = e < C * C ? Rs[e] : make_double2(0.0, 0.0); }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int e = e0 + u * nt + tid, i = e / C, j = e - i * C;
      if (e < C * C && j <= i) A[chol_row_off(i, C) + j] = v[u];
    }
  }
  for (int j = tid; j < C; j += nt) {
    A[chol_row_off(C, C) + j] = make_double2(d[j].x, -d[j].y);
    A[chol_row_off(C + 1, C) + j] = make_double2((j & 1) ? -1.0 - (double)j / C : 1.0 + (double)j / C, 0.0);   // x0 is real
  }
  if (tid == 0) s_bad = 0;
  __syncthreads();
  // (1) the diagonal block of the panel at kp (rows kp .. kp+pw-1), inside warp 0: lane < 28 owns one entry (bi, bj),
  // 1 <= bj <= bi <= 7, of the block's lower triangle (column 0 is never updated), a step is ONE multiply-add per lane
  auto block_factor = [&](int kp, int pw) {
    int bi = 0, bj = 0;                                          // bj = 0: no entry (lanes 28 .. 31)
    {
      int l = lane, b = 1;
      while (b < NP && l >= NP - b) { l -= NP - b; b++; }
      if (b < NP) { bj = b; bi = b + l; }
    }
    const int oi = chol_row_off(kp + bi < C ? kp + bi : 0, C), oj = chol_row_off(kp + bj < C ? kp + bj : 0, C);
    bool bad = false;
    for (int p = 0; p < pw; p++) {
      const int k = kp + p;
      const double Dk = A[chol_row_off(k, C) + k].x;
      if (!(Dk > dThreshold) || !isfinite(Dk)) { bad = true; break; }            // same value in every lane
      if (bj > p && bi < pw) {
        const double rD = 1.0 / Dk;
        const double2 ci = A[oi + k], cj = A[oj + k];
        const double2 f = make_double2(ci.x * rD, ci.y * rD);
        double2 a = A[oi + kp + bj];
        a.x -= f.x * cj.x + f.y * cj.y;                                          // f conj(c_jk)
        a.y -= f.y * cj.x - f.x * cj.y;
        A[oi + kp + bj] = a;
      }
      __syncwarp();
    }
    if (bad) {
      if (lane == 0) s_bad = 1;
      return;
    }
    if (lane < pw) Dinv[kp + lane] = 1.0 / A[chol_row_off(kp + lane, C) + kp + lane].x;
    __syncwarp();
    // conj(L_pq) = conj(c_pq) / D_q for q < p, zero elsewhere (the rows below read the whole table)
    for (int e = lane; e < NP * NP; e += 32) {
      const int pp = e / NP, q = e % NP;
      double2 v = make_double2(0.0, 0.0);
      if (q < pp && pp < pw) { const double2 c = A[chol_row_off(kp + pp, C) + kp + q]; const double r = Dinv[kp + q]; v = make_double2(c.x * r, -c.y * r); }
      Bm[e] = v;
    }
  };
  // (2) rows below the block, one thread per row: c_ip -= sum_{q<p} c_iq conj(L_pq); multipliers W_ip = conj(c_ip) / D_p
  auto rows_below = [&](int kp, int pw) {
    const int i = kp + pw + tid;
    if (i >= NR) return;
    const int oi = chol_row_off(i, C);
    double2 c[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) c[p] = p < pw ? A[oi + kp + p] : make_double2(0.0, 0.0);
#pragma unroll
    for (int p = 1; p < NP; p++) {
#pragma unroll
      for (int q = 0; q < p; q++) {
        const double2 l = Bm[p * NP + q];
        c[p].x -= c[q].x * l.x - c[q].y * l.y;
        c[p].y -= c[q].x * l.y + c[q].y * l.x;
      }
    }
#pragma unroll
    for (int p = 0; p < NP; p++) {
      if (p < pw) {
        A[oi + kp + p] = c[p];
        const double r = Dinv[kp + p];
        Wm[i * NP + p] = make_double2(c[p].x * r, -c[p].y * r);
      } else {
        Wm[i * NP + p] = make_double2(0.0, 0.0);
      }
    }
  };
  // (3) rank-pw update of the rows from `rlo` on, columns [jlo, jhi) (clipped to the row's lower triangle): lanes across the
  // rows, the `nwp` participating warps (this one is number `wp`) across blocks of four columns -- the multipliers of a column
  // are a broadcast load, a row's panel entries stay in registers, eight independent accumulation chains per lane
  auto trailing = [&](int kp, int pw, int rlo, int jlo, int jhi, int wp, int nwp) {
    for (int i = rlo + lane; i < NR; i += 32) {
      const int jtop = i < C ? i : C - 1, jmax = jtop < jhi - 1 ? jtop : jhi - 1;
      const int oi = chol_row_off(i, C);
      double2 c[NP];
#pragma unroll
      for (int p = 0; p < NP; p++) c[p] = p < pw ? A[oi + kp + p] : make_double2(0.0, 0.0);
      for (int j0 = jlo + 4 * wp; j0 <= jmax; j0 += 4 * nwp) {
        double2 a[4];
        const double2* wj[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int j = j0 + q <= jmax ? j0 + q : jmax;          // columns past the clip run on column jmax and are not stored
          a[q] = A[oi + j];
          wj[q] = Wm + j * NP;
        }
#pragma unroll
        for (int p = 0; p < NP; p++) {
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const double2 wv = wj[q][p];
            a[q].x -= c[p].x * wv.x - c[p].y * wv.y;
            a[q].y -= c[p].x * wv.y + c[p].y * wv.x;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) if (j0 + q <= jmax) A[oi + j0 + q] = a[q];
      }
    }
  };
  // Panels with a look-ahead of one: the columns of the NEXT panel are updated first (3a), then warp 0 factors the next
  // diagonal block while the other warps finish the trailing update (3b) -- the serial 8 x 8 block was half of the stall
  // samples when everybody waited for it -- then the rows below the next block (2).  Three CTA barriers per panel.
  {
    const int pw0 = C < NP ? C : NP;
    if (warp == 0) block_factor(0, pw0);
    __syncthreads();
    if (!s_bad) rows_below(0, pw0);
    __syncthreads();
  }
  for (int kp = 0; kp < C && !s_bad; kp += NP) {
    const int pw = C - kp < NP ? C - kp : NP, r0 = kp + pw;
    const int pwn = r0 >= C ? 0 : (C - r0 < NP ? C - r0 : NP);   // width of the next panel
    const int nwarp = nt >> 5;
    if (pwn == 0) {
      // last panel: only the right-hand-side rows are left below it, and they have no columns past C
      break;
    }
    // (3a) the next panel's columns of every row below this panel
    trailing(kp, pw, r0, r0, r0 + pwn, warp, nwarp);
    __syncthreads();
    // (3b) the rest of the trailing matrix (warps 1 ..) | (1) the next diagonal block (warp 0)
    if (warp == 0) block_factor(r0, pwn);
    else trailing(kp, pw, r0 + pwn, r0 + pwn, C, warp - 1, nwarp - 1);
    __syncthreads();
    if (s_bad) break;
    // (2) rows below the next block
    rows_below(r0, pwn);
    __syncthreads();
  }
  const bool bad = s_bad != 0;
  if (!bad && warp == 0) {
    // ---- back substitution L^H t = D^-1 z for z = conj(row C) (from d) and conj(row C+1) (from x0); lane owns rows lane, lane+32
    const int od = chol_row_off(C, C), ox = chol_row_off(C + 1, C);
    double2 rt[2], ru[2];
    double di[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int i = lane + 32 * q;
      di[q] = i < C ? Dinv[i] : 0.0;
      const double2 zd = i < C ? A[od + i] : make_double2(0.0, 0.0), zx = i < C ? A[ox + i] : make_double2(0.0, 0.0);
      rt[q] = make_double2(zd.x * di[q], -zd.y * di[q]);
      ru[q] = make_double2(zx.x * di[q], -zx.y * di[q]);
    }
    for (int r = C - 1; r > 0; r--) {
      const int src = r & 31;
      double2 tr = r >= 32 ? rt[1] : rt[0], ur = r >= 32 ? ru[1] : ru[0];
      tr.x = shfl_d(tr.x, src); tr.y = shfl_d(tr.y, src); ur.x = shfl_d(ur.x, src); ur.y = shfl_d(ur.y, src);
      const int orow = chol_row_off(r, C);
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int i = lane + 32 * q;
        if (i < r) {
          const double2 c = A[orow + i];                         // c_ri = L_ri D_i;  conj(L_ri) = conj(c_ri) / D_i
          const double lx = c.x * di[q], ly = -c.y * di[q];
          rt[q].x -= lx * tr.x - ly * tr.y; rt[q].y -= lx * tr.y + ly * tr.x;
          ru[q].x -= lx * ur.x - ly * ur.y; ru[q].y -= lx * ur.y + ly * ur.x;
        }
      }
    }
    // lam = t^H d, |t|^2, |d|^2, |u|^2, |x0|^2
    double2 lam = make_double2(0.0, 0.0);
    double nt2 = 0.0, nd2 = 0.0, nu2 = 0.0, nx2 = 0.0;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int c = lane + 32 * q;
      if (c < C) {
        const double2 dc = d[c];
        lam.x += rt[q].x * dc.x + rt[q].y * dc.y; lam.y += rt[q].x * dc.y - rt[q].y * dc.x;
        nt2 += rt[q].x * rt[q].x + rt[q].y * rt[q].y; nd2 += dc.x * dc.x + dc.y * dc.y;
        nu2 += ru[q].x * ru[q].x + ru[q].y * ru[q].y;
        const double x0 = (c & 1) ? -1.0 - (double)c / C : 1.0 + (double)c / C;
        nx2 += x0 * x0;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lam.x += shfl_d(lam.x, lane ^ o); lam.y += shfl_d(lam.y, lane ^ o);
      nt2 += shfl_d(nt2, lane ^ o); nd2 += shfl_d(nd2, lane ^ o); nu2 += shfl_d(nu2, lane ^ o); nx2 += shfl_d(nx2, lane ^ o);
    }
    // sigma_min <= |x| / |R^-1 x|: reject when either probe puts it below the threshold
    const double thr2 = dThreshold * dThreshold;
    const bool rej = dThreshold > 0.0 && (!(nx2 >= thr2 * nu2) || !(nd2 >= thr2 * nt2) || !isfinite(nu2) || !isfinite(nt2));
    if (rej || !isfinite(lam.x) || !isfinite(lam.y) || (lam.x == 0.0 && lam.y == 0.0)) {
      if (lane == 0) s_bad = 1;
    } else {
      const double2 nrm = make_double2(lam.x * C, lam.y * C);
#pragma unroll
      for (int q = 0; q < 2; q++) { const int c = lane + 32 * q; if (c < C) ws[c] = zdiv(rt[q], nrm); }
      if (lane == 0) fallback[s] = 0;
    }
  }
  __syncthreads();
  if (s_bad) {
    // identity fallback (beamformer.cc:2425-2427): t = d
    if (tid == 0) {
      double lr = 0.0;
      for (int c = 0; c < C; c++) lr += d[c].x * d[c].x + d[c].y * d[c].y;
      for (int c = 0; c < C; c++) ws[c] = make_double2(d[c].x / (lr * C), d[c].y / (lr * C));
      fallback[s] = 1;
    }
  }
}

cudaError_t launch_mvdr_chol(const double2* Rn, const double2* d, double2* w, int* fallback, int B, int C, double dThreshold,
                             cudaStream_t st, int n) {
  if (C > 64) return cudaErrorInvalidValue;
  const int NR = C + 2;
  const size_t smem = ((size_t)(C * (C + 1)) / 2 + 2 * C + (size_t)NR * BTK_CHOL_PANEL + BTK_CHOL_PANEL * BTK_CHOL_PANEL) * sizeof(double2) +
                      (size_t)C * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(btk_mvdr_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  btk_mvdr_chol_kernel<<<B * n, 128, smem, st>>>(Rn, d, w, fallback, C, dThreshold, B);
  return cudaGetLastError();
}

cudaError_t launch_mvdr_solve(const double2* Rn, const double2* d, double2* w, int* fallback, int B, int C,
                              double dThreshold, cudaStream_t st, int n) {
  const size_t smem = (size_t)C * (C + 1) * sizeof(double2);
  cudaError_t e = cudaFuncSetAttribute(btk_mvdr_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  btk_mvdr_solve_kernel<<<B * n, C > 32 ? 256 : 128, smem, st>>>(Rn, d, w, fallback, C, dThreshold, B);
  return cudaGetLastError();
}

}  // namespace btk
