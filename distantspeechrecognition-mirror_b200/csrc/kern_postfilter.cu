// kern_postfilter.cu -- Zelinski post-filter between the beamformer and the synthesis bank
// (postfilter/postfilter.cc:30-222 ZelinskiFilter_f / ZelinskiFilter, :428-500 ZelinskiPostFilter::next; used by both
// shipped drivers, src/superdirectiveBeamformer.cc:164,200 and src/beamformerDS.cc:154,183).
//
// Reference, per frame n and bin s (0..M/2):  y_i = conj(ta_i) x_i  (time alignment with the array manifold),
//     Phi_ij(n) = a_n Phi_ij(n-1) + (1 - a_n) y_i conj(y_j)   for every pair i < j,      (calcCSD, :8-21)
//     Psi_i(n)  = a_n Psi_i(n-1)  + (1 - a_n) |y_i|^2,
//     W = clamp( num / sum_i Psi_i * 2 / (C - 1), 1e-4, 1 ),  num = |sum_{i<j} Phi_ij| (ABS) or max(Re sum, 0) (REAL),
// with a_n = alpha for n >= 2 and 0 for the first two frames (alpha is applied only once _frameX > 0, :466-469), and
// the beamformer output is multiplied by W from frame minFrames + 1 on (:474-479).
//
// The recursions are linear and share a_n, so the sum over pairs obeys the same recursion: only
//     S(n) = a_n S(n-1) + (1 - a_n) s(n),  s(n) = sum_{i<j} y_i conj(y_j)   and   P(n), p(n) = sum_i |y_i|^2
// are carried -- O(C) per (frame, bin) with a running prefix sum instead of O(C^2) state per bin.
//   kernel 1 (frame- and bin-parallel): beamformer output Y, s, p from the stored snapshots (one coalesced pass);
//   kernels 2-4: the first-order recursion over frames as a three-phase segmented scan (affine maps compose), W, Y *= W.
#include "launch.h"
#include "snap_tile.cuh"

namespace btk {

// A lane group per (frame, bin) item, channels across the lanes in rows of GS (snap_tile.cuh).  The ordered pair sum
//   s = sum_j (sum_{i<j} y_i) conj(y_j)
// needs the prefix over the channel index: within a row of GS channels an exclusive group scan, across rows the
// running total of the earlier rows.
template <int GS>
__global__ void __launch_bounds__(SNAP_THREADS) btk_beamform_zelinski_kernel(const cf* __restrict__ snap, const cf* __restrict__ w,
                                                                            const cf* __restrict__ ta, cf* __restrict__ Y,
                                                                            float4* __restrict__ stat, long long FB, int B, int C) {
  const int lg = threadIdx.x % GS;
  const long long g0 = ((long long)blockIdx.x * SNAP_THREADS + threadIdx.x) / GS;
  const long long gstride = (long long)gridDim.x * (SNAP_THREADS / GS);
  const long long n_it = (FB + gstride - 1) / gstride;
  const int rows = (C + GS - 1) / GS;
  for (long long it = 0; it < n_it; it++) {
    const long long idx = g0 + it * gstride;
    const bool live = idx < FB;
    const int s = live ? (int)(idx % B) : 0;
    const cf* x = snap + (live ? idx : 0) * C;
    const cf* ws = w + (long long)s * C;
    const cf* ts = ta + (long long)s * C;
    float yr = 0.f, yi = 0.f;            // beamformer output  sum_c conj(w_c) x_c   (beamformer.cc:1181-1188)
    float tr = 0.f, ti = 0.f;            // sum of the aligned channels of the earlier rows
    float sr = 0.f, si = 0.f, p = 0.f;
    for (int row = 0; row < rows; row++) {
      const int c = row * GS + lg;
      float ar = 0.f, ai = 0.f;
      if (live && c < C) {
        const cf a = __ldg(ws + c), t = __ldg(ts + c), b = x[c];
        yr = fmaf(a.x, b.x, yr); yr = fmaf(a.y, b.y, yr);
        yi = fmaf(a.x, b.y, yi); yi = fmaf(-a.y, b.x, yi);
        ar = t.x * b.x + t.y * b.y; ai = t.x * b.y - t.y * b.x;      // y_c = conj(ta_c) x_c
      }
      const float pr = tr + group_exscan<GS>(ar, lg), pi = ti + group_exscan<GS>(ai, lg);   // sum_{i<c} y_i
      sr += pr * ar + pi * ai;           // (sum_{i<c} y_i) conj(y_c)
      si += pi * ar - pr * ai;
      p = fmaf(ar, ar, fmaf(ai, ai, p));
      tr += group_sum<GS>(ar);
      ti += group_sum<GS>(ai);
    }
    yr = group_sum<GS>(yr); yi = group_sum<GS>(yi);
    sr = group_sum<GS>(sr); si = group_sum<GS>(si); p = group_sum<GS>(p);
    if (live && lg == 0) {
      Y[idx] = mk(yr, yi);
      stat[idx] = make_float4(sr, si, p, 0.f);
    }
  }
}

// ---- the recursion over frames as a three-phase segmented scan (affine maps compose) ---------------------------------
// segments of ZEL_LEN frames; thread = (bin, segment), lanes across 32 consecutive bins (coalesced).
//   phase 1: every segment from a zero state -> its end state and the product of its coefficients
//   phase 2: per bin, fold the segments in order -> the state entering every segment   (F / ZEL_LEN dependent steps)
//   phase 3: every segment again from its entering state -> gains W, Y *= W
#define ZEL_LEN 64
#define ZEL_WARPS 8

__device__ __forceinline__ double zel_alpha(long long n, double alpha) { return n >= 2 ? alpha : 0.0; }   // postfilter.cc:466-469

__global__ void __launch_bounds__(32 * ZEL_WARPS) btk_zelinski_seg_kernel(const float4* __restrict__ stat, double4* __restrict__ seg,
                                                                         long long F, int B, int nseg, double alpha) {
  const int s = blockIdx.x * 32 + (threadIdx.x & 31), g = blockIdx.y * ZEL_WARPS + (threadIdx.x >> 5);
  if (s >= B || g >= nseg) return;
  const long long n0 = (long long)g * ZEL_LEN, n1 = n0 + ZEL_LEN < F ? n0 + ZEL_LEN : F;
  double Sr = 0.0, Si = 0.0, P = 0.0, A = 1.0;
  for (long long nb = n0; nb < n1; nb += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = nb + u < n1 ? stat[(nb + u) * B + s] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (nb + u >= n1) break;
      const double a = zel_alpha(nb + u, alpha);
      Sr = a * Sr + (1.0 - a) * (double)v[u].x;
      Si = a * Si + (1.0 - a) * (double)v[u].y;
      P = a * P + (1.0 - a) * (double)v[u].z;
      A *= a;
    }
  }
  seg[(long long)g * B + s] = make_double4(Sr, Si, P, A);
}

// in place: seg[g] <- state ENTERING segment g
__global__ void btk_zelinski_fold_kernel(double4* __restrict__ seg, int B, int nseg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  double Sr = 0.0, Si = 0.0, P = 0.0;
  for (int g0 = 0; g0 < nseg; g0 += 8) {           // the loads do not depend on the state: eight in flight
    double4 e[8];
#pragma unroll
    for (int u = 0; u < 8; u++) e[u] = g0 + u < nseg ? seg[(long long)(g0 + u) * B + s] : make_double4(0.0, 0.0, 0.0, 1.0);
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (g0 + u >= nseg) break;
      seg[(long long)(g0 + u) * B + s] = make_double4(Sr, Si, P, 0.0);
      Sr = e[u].w * Sr + e[u].x; Si = e[u].w * Si + e[u].y; P = e[u].w * P + e[u].z;
    }
  }
}

__global__ void __launch_bounds__(32 * ZEL_WARPS) btk_zelinski_apply_kernel(const float4* __restrict__ stat, const double4* __restrict__ seg,
                                                                           cf* __restrict__ Y, float* __restrict__ Wout, long long F,
                                                                           int B, int C, int nseg, double alpha, int type,
                                                                           int min_frames) {
  const int s = blockIdx.x * 32 + (threadIdx.x & 31), g = blockIdx.y * ZEL_WARPS + (threadIdx.x >> 5);
  if (s >= B || g >= nseg) return;
  const long long n0 = (long long)g * ZEL_LEN, n1 = n0 + ZEL_LEN < F ? n0 + ZEL_LEN : F;
  const double4 e = seg[(long long)g * B + s];
  double Sr = e.x, Si = e.y, P = e.z;
  const double scale = 2.0 / ((double)C - 1.0);             // 2 / (nChan - 1), postfilter.cc:121
  for (long long nb = n0; nb < n1; nb += 4) {
    float4 v[4];
    cf y[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const bool ok = nb + u < n1;
      v[u] = ok ? stat[(nb + u) * B + s] : make_float4(0.f, 0.f, 0.f, 0.f);
      y[u] = ok ? Y[(nb + u) * B + s] : mk(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long n = nb + u;
      if (n >= n1) break;
      const double a = zel_alpha(n, alpha);
      Sr = a * Sr + (1.0 - a) * (double)v[u].x;
      Si = a * Si + (1.0 - a) * (double)v[u].y;
      P = a * P + (1.0 - a) * (double)v[u].z;
      const bool applied = n > min_frames && type != 0;     // frames 0 .. minFrames only update the densities, and
                                                            // their (unused) gain follows the |.| branch (pfType = 0)
      double num;
      if (applied && (type & 1)) num = Sr < 0.0 ? 0.0 : Sr; // TYPE_ZELINSKI1_REAL
      else num = sqrt(Sr * Sr + Si * Si);
      double W = (num / P) * scale;
      if (W >= 1.0) W = 1.0;
      if (W < 1.0e-4) W = 1.0e-4;                            // SPECTRAL_FLOOR
      if (Wout) Wout[n * B + s] = (float)W;
      if (applied) Y[n * B + s] = mk(y[u].x * (float)W, y[u].y * (float)W);
    }
  }
}

cudaError_t launch_beamform_zelinski(const cf* snap, const cf* w, const cf* ta, cf* Y, float4* stat, float* Wout, long long F,
                                     int B, int C, double alpha, int type, int min_frames, cudaStream_t st) {
  const long long FB = F * B;
  if (FB == 0) return cudaSuccess;
  if (C < 2) return cudaErrorInvalidValue;                  // jdimension_error in the reference (postfilter.cc:62-65)
  const int grid = snap_grid(FB, C);
  switch (snap_group_size(C)) {
    case 2: btk_beamform_zelinski_kernel<2><<<grid, SNAP_THREADS, 0, st>>>(snap, w, ta, Y, stat, FB, B, C); break;
    case 4: btk_beamform_zelinski_kernel<4><<<grid, SNAP_THREADS, 0, st>>>(snap, w, ta, Y, stat, FB, B, C); break;
    case 8: btk_beamform_zelinski_kernel<8><<<grid, SNAP_THREADS, 0, st>>>(snap, w, ta, Y, stat, FB, B, C); break;
    case 16: btk_beamform_zelinski_kernel<16><<<grid, SNAP_THREADS, 0, st>>>(snap, w, ta, Y, stat, FB, B, C); break;
    default: btk_beamform_zelinski_kernel<32><<<grid, SNAP_THREADS, 0, st>>>(snap, w, ta, Y, stat, FB, B, C); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // the scan scratch (one double4 per bin and segment) follows the statistics in the caller's buffer
  const int nseg = (int)((F + ZEL_LEN - 1) / ZEL_LEN);
  double4* seg = reinterpret_cast<double4*>(stat + ((FB + 1) & ~1LL));
  const dim3 sgrid((B + 31) / 32, (nseg + ZEL_WARPS - 1) / ZEL_WARPS);
  btk_zelinski_seg_kernel<<<sgrid, 32 * ZEL_WARPS, 0, st>>>(stat, seg, F, B, nseg, alpha);
  btk_zelinski_fold_kernel<<<(B + 127) / 128, 128, 0, st>>>(seg, B, nseg);
  btk_zelinski_apply_kernel<<<sgrid, 32 * ZEL_WARPS, 0, st>>>(stat, seg, Y, Wout, F, B, C, nseg, alpha, type, min_frames);
  return cudaGetLastError();
}

}  // namespace btk
