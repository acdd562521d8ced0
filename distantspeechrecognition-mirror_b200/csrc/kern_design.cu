// kern_design.cu -- de Haan prototype design for the oversampled DFT filter bank, all in fp64 on the device:
//   analysis   h = pinv(A + C) b        AnalysisOversampledDFTDesign   (modulated/prototypeDesign.cc:223-272, 640-712)
//   synthesis  g = pinv(E + v P) f      SynthesisOversampledDFTDesign  (modulated/prototypeDesign.cc:836-901)
// with the truncated pseudo-inverse of the reference: singular values below tolerance * s_max are dropped.
//
// The L x L matrices (L = M m, up to 4096) are symmetric, so their SVD is computed by a parallel one-sided Jacobi
// (Hestenes) iteration: columns of G = K V are rotated in pairs until they are mutually orthogonal; then
// s_j = |g_j|, u_j = g_j / s_j and  x = sum_j v_j (g_j . rhs) / s_j^2.  One step rotates L/2 disjoint column pairs
// (round-robin tournament ordering), one CTA per pair; a sweep is L-1 steps.  G and V (2 x 8 MB at L = 1024) live in
// L2 for the whole iteration.  No cuSOLVER / cuBLAS.
#include <math.h>
#include <stdio.h>

#include <vector>

#include "launch.h"

namespace btk {

#define DSG_THREADS 256

// ---- matrix builders (column-major; the matrices are symmetric) -------------------------------------------------
__device__ __forceinline__ double dsg_factor(long long d, int D) { return (d % D == 0) ? (double)(D - 1) : -1.0; }

// K = A + C, rhs = b; A and C are also kept for calcError()
__global__ void dsg_analysis_matrices(double* __restrict__ K, double* __restrict__ A, double* __restrict__ C, double* __restrict__ b,
                                      int L, int D, double wp, int tau) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)L * L) return;
  const int m = (int)(idx % L), n = (int)(idx / L);
  const int d = n - m;
  const double a = d == 0 ? 1.0 : sin(wp * d) / (wp * d);                       // :232-236
  const double f = dsg_factor(d, D);
  const double c = d == 0 ? f / D : f * sin(M_PI * d / D) / (M_PI * D * d);       // :256-266
  A[idx] = a; C[idx] = c; K[idx] = a + c;
  if (n == 0) {
    const int t = tau - m;
    b[m] = t == 0 ? 1.0 : sin(wp * t) / (wp * t);                               // :239-243
  }
}

// rr[d] = sum_j h[j] h[j + d], d = 0..L-1
__global__ void dsg_autocorr(const double* __restrict__ h, double* __restrict__ rr, int L) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= L) return;
  double s = 0;
  for (int j = 0; j + d < L; j++) s += h[j] * h[j + d];
  rr[d] = s;
}

// K = E + v P, rhs = f; E and P kept for calcError()   (:836-871)
__global__ void dsg_synthesis_matrices(double* __restrict__ K, double* __restrict__ E, double* __restrict__ P, double* __restrict__ f,
                                       const double* __restrict__ h, const double* __restrict__ rr, int L, int M, int mm, int D,
                                       double v, int tau) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)L * L) return;
  const int m = (int)(idx % L), n = (int)(idx / L);
  double e = 0;
  for (int k = 0; k <= 2 * mm; k++) {
    const int a = k * M - m, b = k * M - n;
    if (a < 0 || a > L - 1 || b < 0 || b > L - 1) continue;
    e += h[a] * h[b];
  }
  const int R = M / D;
  e *= (double)(R * R);
  const int d = m - n;
  const double p = dsg_factor(d, D) * rr[d < 0 ? -d : d] * ((double)M / ((double)D * (double)D));
  E[idx] = e; P[idx] = p; K[idx] = e + v * p;
  if (n == 0) {
    const int t = 2 * tau - m;
    f[m] = (t < 0 || t > L - 1) ? 0.0 : h[t] * ((double)M / (M_PI * D));
  }
}

__global__ void dsg_identity(double* __restrict__ V, int L) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)L * L) return;
  V[idx] = (idx % L == idx / L) ? 1.0 : 0.0;
}

// ---- one Jacobi step: pair i of round `step` of the round-robin tournament over L columns --------------------------
__device__ __forceinline__ double dsg_block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < DSG_THREADS / 32; w++) s += red[w];
  return s;
}

__global__ void __launch_bounds__(DSG_THREADS) dsg_jacobi_step(double* __restrict__ G, double* __restrict__ V, int L, int step,
                                                              unsigned long long* __restrict__ off_bits) {
  __shared__ double red[DSG_THREADS / 32];
  const int i = blockIdx.x, n1 = L - 1;
  int p, q;
  if (i == 0) { p = step % n1; q = n1; }
  else { p = (step + i) % n1; q = (step - i + n1) % n1; }
  if (p > q) { const int t = p; p = q; q = t; }
  double* gp = G + (size_t)p * L;
  double* gq = G + (size_t)q * L;
  double a = 0, b = 0, g = 0;
  for (int r = threadIdx.x; r < L; r += DSG_THREADS) { const double x = gp[r], y = gq[r]; a += x * x; b += y * y; g += x * y; }
  a = dsg_block_sum(a, red); b = dsg_block_sum(b, red); g = dsg_block_sum(g, red);
  if (g == 0.0 || a == 0.0 || b == 0.0) return;
  const double rel = fabs(g) / sqrt(a * b);
  if (threadIdx.x == 0) atomicMax(off_bits, (unsigned long long)__double_as_longlong(rel));   // rel >= 0: bit order = value order
  if (rel < 1e-15) return;
  const double zeta = (b - a) / (2.0 * g);
  const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
  for (int r = threadIdx.x; r < L; r += DSG_THREADS) { const double x = gp[r], y = gq[r]; gp[r] = c * x - s * y; gq[r] = s * x + c * y; }
  double* vp = V + (size_t)p * L;
  double* vq = V + (size_t)q * L;
  for (int r = threadIdx.x; r < L; r += DSG_THREADS) { const double x = vp[r], y = vq[r]; vp[r] = c * x - s * y; vq[r] = s * x + c * y; }
}

// sig2[j] = |g_j|^2, gb[j] = g_j . rhs
__global__ void __launch_bounds__(DSG_THREADS) dsg_column_stats(const double* __restrict__ G, const double* __restrict__ rhs,
                                                               double* __restrict__ sig2, double* __restrict__ gb, int L) {
  __shared__ double red[DSG_THREADS / 32];
  const double* gj = G + (size_t)blockIdx.x * L;
  double a = 0, b = 0;
  for (int r = threadIdx.x; r < L; r += DSG_THREADS) { const double x = gj[r]; a += x * x; b += x * rhs[r]; }
  a = dsg_block_sum(a, red); b = dsg_block_sum(b, red);
  if (threadIdx.x == 0) { sig2[blockIdx.x] = a; gb[blockIdx.x] = b; }
}

// x[i] = sum_j V[i][j] c[j]   (V column-major: V[i + j L])
__global__ void dsg_combine(const double* __restrict__ V, const double* __restrict__ c, double* __restrict__ x, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  double s = 0;
  for (int j = 0; j < L; j++) s += V[i + (size_t)j * L] * c[j];
  x[i] = s;
}

// out[0] = x^T Mat x, out[1] = x^T y   (Mat symmetric, column-major)
__global__ void __launch_bounds__(DSG_THREADS) dsg_quadratic(const double* __restrict__ Mat, const double* __restrict__ x,
                                                            const double* __restrict__ y, double* __restrict__ out, int L) {
  __shared__ double red[DSG_THREADS / 32];
  double q = 0, d = 0;
  for (int j = threadIdx.x; j < L; j += DSG_THREADS) {
    const double* col = Mat + (size_t)j * L;
    double s = 0;
    for (int i = 0; i < L; i++) s += col[i] * x[i];
    q += s * x[j];
    if (y) d += x[j] * y[j];
  }
  q = dsg_block_sum(q, red); d = dsg_block_sum(d, red);
  if (threadIdx.x == 0) { out[0] = q; out[1] = d; }
}

#define DSG_CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = e__; goto done; } } while (0)

// Truncated pseudo-inverse solve of the symmetric system held in d_K (destroyed): x = pinv(K) rhs.
static cudaError_t dsg_pinv_solve(double* d_K, double* d_V, const double* d_rhs, double* d_x, double* d_tmp /*3L + 2*/, int L,
                                  double tolerance, int* sweeps_out) {
  cudaError_t rc = cudaSuccess;
  unsigned long long* d_off = reinterpret_cast<unsigned long long*>(d_tmp + 3 * (size_t)L);
  std::vector<double> sig2(L), gb(L), c(L);
  const long long LL = (long long)L * L;
  int sweeps = 0;
  dsg_identity<<<(unsigned)((LL + 255) / 256), 256>>>(d_V, L);
  DSG_CK(cudaGetLastError());
  for (sweeps = 0; sweeps < 40; sweeps++) {
    DSG_CK(cudaMemset(d_off, 0, sizeof(unsigned long long)));
    for (int step = 0; step < L - 1; step++) dsg_jacobi_step<<<L / 2, DSG_THREADS>>>(d_K, d_V, L, step, d_off);
    DSG_CK(cudaGetLastError());
    unsigned long long bits = 0;
    DSG_CK(cudaMemcpy(&bits, d_off, sizeof bits, cudaMemcpyDeviceToHost));
    double off;
    memcpy(&off, &bits, sizeof off);
    if (off < 1e-15) { sweeps++; break; }
  }
  dsg_column_stats<<<L, DSG_THREADS>>>(d_K, d_rhs, d_tmp, d_tmp + L, L);
  DSG_CK(cudaGetLastError());
  DSG_CK(cudaMemcpy(sig2.data(), d_tmp, L * sizeof(double), cudaMemcpyDeviceToHost));
  DSG_CK(cudaMemcpy(gb.data(), d_tmp + L, L * sizeof(double), cudaMemcpyDeviceToHost));
  {
    // the reference divides U^T rhs by every singular value with s_n / s_0 >= tolerance and zeroes the rest (:699-709)
    double smax = 0;
    for (int j = 0; j < L; j++) if (sig2[j] > smax) smax = sig2[j];
    smax = sqrt(smax);
    for (int j = 0; j < L; j++) {
      const double s = sqrt(sig2[j]);
      c[j] = (smax > 0 && s / smax >= tolerance && s > 0) ? gb[j] / sig2[j] : 0.0;
    }
  }
  DSG_CK(cudaMemcpy(d_tmp + 2 * (size_t)L, c.data(), L * sizeof(double), cudaMemcpyHostToDevice));
  dsg_combine<<<(L + 127) / 128, 128>>>(d_V, d_tmp + 2 * (size_t)L, d_x, L);
  DSG_CK(cudaGetLastError());
done:
  if (sweeps_out) *sweeps_out = sweeps;
  return rc;
}

// kind 0: analysis (h_in unused), kind 1: synthesis from h_in.  proto_out [L]; err [3] (calcError) or NULL.
cudaError_t design_prototype(int kind, const double* h_in, int M, int m, int r, double v, double wp_factor, int tau, double tolerance,
                             double* proto_out, double* err, int* sweeps_out) {
  const int L = M * m, D = M >> r;
  if (tau < 0) tau = L / 2;                                                     // :203
  const long long LL = (long long)L * L;
  cudaError_t rc = cudaSuccess;
  double *d_K = 0, *d_V = 0, *d_A = 0, *d_B = 0, *d_rhs = 0, *d_x = 0, *d_tmp = 0, *d_h = 0, *d_rr = 0;
  double q[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  DSG_CK(cudaMalloc(&d_K, LL * sizeof(double)));
  DSG_CK(cudaMalloc(&d_V, LL * sizeof(double)));
  DSG_CK(cudaMalloc(&d_A, LL * sizeof(double)));
  DSG_CK(cudaMalloc(&d_B, LL * sizeof(double)));
  DSG_CK(cudaMalloc(&d_rhs, L * sizeof(double)));
  DSG_CK(cudaMalloc(&d_x, L * sizeof(double)));
  DSG_CK(cudaMalloc(&d_tmp, (3 * (size_t)L + 16) * sizeof(double)));
  DSG_CK(cudaMalloc(&d_h, L * sizeof(double)));
  DSG_CK(cudaMalloc(&d_rr, L * sizeof(double)));
  if (kind == 0) {
    dsg_analysis_matrices<<<(unsigned)((LL + 255) / 256), 256>>>(d_K, d_A, d_B, d_rhs, L, D, M_PI / (wp_factor * M), tau);
  } else {
    DSG_CK(cudaMemcpy(d_h, h_in, L * sizeof(double), cudaMemcpyHostToDevice));
    dsg_autocorr<<<(L + 127) / 128, 128>>>(d_h, d_rr, L);
    dsg_synthesis_matrices<<<(unsigned)((LL + 255) / 256), 256>>>(d_K, d_A, d_B, d_rhs, d_h, d_rr, L, M, m, D, v, tau);
  }
  DSG_CK(cudaGetLastError());
  DSG_CK(dsg_pinv_solve(d_K, d_V, d_rhs, d_x, d_tmp, L, tolerance, sweeps_out));
  DSG_CK(cudaMemcpy(proto_out, d_x, L * sizeof(double), cudaMemcpyDeviceToHost));
  if (err) {
    double* d_out = d_tmp;
    if (kind == 0) {
      // eps_p = 10 log10(h^T A h - 2 h^T b + 1), eps_i = 10 log10(h^T C h)   (:724-766)
      dsg_quadratic<<<1, DSG_THREADS>>>(d_A, d_x, d_rhs, d_out, L);
      dsg_quadratic<<<1, DSG_THREADS>>>(d_B, d_x, (const double*)0, d_out + 2, L);
      DSG_CK(cudaMemcpy(q, d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost));
      err[0] = 10.0 * log10(q[0][0] - 2.0 * q[0][1] + 1.0);
      err[1] = 10.0 * log10(q[1][0]);
      err[2] = err[0] + err[1];
    } else {
      // eps_t = 10 log10(g^T E g - 2 g^T b + 1) with the base-class b, which this design never fills (all zero:
      // SynthesisOversampledDFTDesign::design does not call _calculateAb, :800-812, 903-933); eps_r = 10 log10(g^T P g)
      dsg_quadratic<<<1, DSG_THREADS>>>(d_A, d_x, (const double*)0, d_out, L);
      dsg_quadratic<<<1, DSG_THREADS>>>(d_B, d_x, (const double*)0, d_out + 2, L);
      DSG_CK(cudaMemcpy(q, d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost));
      err[0] = 10.0 * log10(q[0][0] + 1.0);
      err[1] = 10.0 * log10(q[1][0]);
      err[2] = err[0] + v * err[1];
    }
  }
done:
  cudaFree(d_K); cudaFree(d_V); cudaFree(d_A); cudaFree(d_B); cudaFree(d_rhs); cudaFree(d_x); cudaFree(d_tmp); cudaFree(d_h); cudaFree(d_rr);
  return rc;
}

}  // namespace btk
