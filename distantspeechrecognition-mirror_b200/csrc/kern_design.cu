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
#include <stdlib.h>

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


// =====================================================================================================================
// Nyquist(M)-constrained designs (modulated/prototypeDesign.cc:955-1119): AnalysisNyquistMDesign, SynthesisNyquistMDesign.
//   constraints  Hc^T x = c0  (analysis: h[n M] = delta(n - m/2) / M, :982-989; synthesis: the 2 m total-response
//                equations H^T g = c0 with H[k][n] = h[n M - k], c0[m] = D / M, :1073-1089)
//   cond([Hc^T; Q]) < 1 / tolerance -> "alternate solution 4" (_solveNonSingular, :361-470): minimise x^T Q x on the
//                constraint set: x = x_pt - B pinv(B^T Q B) B^T Q x_pt, B = null space of Hc^T, x_pt = pinv(Hc^T) c0
//   otherwise   -> "alternate solution 3" (_solveSingular, :481-577): constraints K = [Hc Q]; x = x_pt + B pinv'(B^T A B)
//                B^T (b - A x_pt) over the numerical null space B of K^T; pinv' divides the components above the
//                tolerance and leaves the others unscaled (:556-558)
// Every decomposition is the same one-sided Jacobi iteration as above, run on rectangular matrices: the rotations
// accumulate in an orthogonal V whatever the rank, so the columns of V that belong to vanishing singular values ARE an
// orthonormal basis of the null space (the reference gets them from GSL's full U of the transposed problem, :280-287).
// =====================================================================================================================

// round-robin pair of block i in round `step` over n columns (n even, or odd with one bye per round)
__device__ __forceinline__ bool dsg_pair(int n, int step, int i, int& p, int& q) {
  const int ne = (n + 1) & ~1, n1 = ne - 1;
  if (i == 0) { p = step % n1; q = n1; }
  else { p = (step + i) % n1; q = (step - i + n1) % n1; }
  if (p > q) { const int t = p; p = q; q = t; }
  return q < n;                                   // the phantom column of an odd n sits out
}

// rectangular variant: G is [rows x n] column-major (leading dimension ldg), V is [n x n]
// floor2: columns whose squared norm is below it have vanished (a wide matrix has n - rows of them): they take no more
// rotations -- two noise columns would otherwise keep an O(1) normalised inner product for ever
__global__ void __launch_bounds__(DSG_THREADS) dsg_jacobi_step_rect(double* __restrict__ G, int rows, int ldg, double* __restrict__ V, int n,
                                                                   int step, unsigned long long* __restrict__ off_bits, double floor2) {
  __shared__ double red[DSG_THREADS / 32];
  int p, q;
  if (!dsg_pair(n, step, blockIdx.x, p, q)) return;
  double* gp = G + (size_t)p * ldg;
  double* gq = G + (size_t)q * ldg;
  double a = 0, b = 0, g = 0;
  for (int r = threadIdx.x; r < rows; r += DSG_THREADS) { const double x = gp[r], y = gq[r]; a += x * x; b += y * y; g += x * y; }
  a = dsg_block_sum(a, red); b = dsg_block_sum(b, red); g = dsg_block_sum(g, red);
  if (g == 0.0 || !(a > floor2) || !(b > floor2)) return;
  const double rel = fabs(g) / sqrt(a * b);
  if (threadIdx.x == 0) atomicMax(off_bits, (unsigned long long)__double_as_longlong(rel));
  if (rel < 1e-15) return;
  const double zeta = (b - a) / (2.0 * g);
  const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
  for (int r = threadIdx.x; r < rows; r += DSG_THREADS) { const double x = gp[r], y = gq[r]; gp[r] = c * x - s * y; gq[r] = s * x + c * y; }
  double* vp = V + (size_t)p * n;
  double* vq = V + (size_t)q * n;
  for (int r = threadIdx.x; r < n; r += DSG_THREADS) { const double x = vp[r], y = vq[r]; vp[r] = c * x - s * y; vq[r] = s * x + c * y; }
}

// sig2[j] = |g_j|^2, gb[j] = g_j . rhs (rhs of `rows` entries), rectangular
__global__ void __launch_bounds__(DSG_THREADS) dsg_column_stats_rect(const double* __restrict__ G, int rows, int ldg, const double* __restrict__ rhs,
                                                                    double* __restrict__ sig2, double* __restrict__ gb) {
  __shared__ double red[DSG_THREADS / 32];
  const double* gj = G + (size_t)blockIdx.x * ldg;
  double a = 0, b = 0;
  for (int r = threadIdx.x; r < rows; r += DSG_THREADS) { const double x = gj[r]; a += x * x; b += x * rhs[r]; }
  a = dsg_block_sum(a, red); b = dsg_block_sum(b, red);
  if (threadIdx.x == 0) { sig2[blockIdx.x] = a; gb[blockIdx.x] = b; }
}

// C[i + j ldc] = sum_k A(i,k) B(k,j); A(i,k) = A[i sai + k sak], B(k,j) = B[k sbk + j sbj]  (any layout / transpose)
__global__ void dsg_gemm(double* __restrict__ C, int ldc, const double* __restrict__ A, long long sai, long long sak,
                         const double* __restrict__ B, long long sbk, long long sbj, int ni, int nj, int nk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= ni || j >= nj) return;
  double s = 0;
  for (int k = 0; k < nk; k++) s += A[i * sai + k * sak] * B[k * sbk + j * sbj];
  C[i + (size_t)j * ldc] = s;
}

// dst[:, j] = src[:, idx[j]]  (columns of length n)
__global__ void dsg_gather_cols(double* __restrict__ dst, const double* __restrict__ src, const int* __restrict__ idx, int n, int ncols) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (r < n && j < ncols) dst[r + (size_t)j * n] = src[r + (size_t)idx[j] * n];
}

// Kt = [Hc^T; Q] column-major [(k + L) x L]: column j = (Hc[j][0..k), Q[:, j])
__global__ void dsg_stack(double* __restrict__ Kt, const double* __restrict__ Hc /*[L][k] row-major*/, const double* __restrict__ Q, int L, int k) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (r >= k + L || j >= L) return;
  Kt[r + (size_t)j * (k + L)] = r < k ? Hc[(size_t)j * k + r] : Q[(r - k) + (size_t)j * L];
}

// One-sided Jacobi SVD of G [rows x n] in place; V [n x n].  Host copies of |g_j|^2 (sig2) and g_j . rhs (gb, rhs may be NULL).
static cudaError_t dsg_svd_rect(double* d_G, int rows, int ldg, double* d_V, int n, const double* d_rhs, double* d_tmp /*2 n + 2*/,
                                std::vector<double>& sig2, std::vector<double>& gb, int* sweeps_total) {
  cudaError_t rc = cudaSuccess;
  unsigned long long* d_off = reinterpret_cast<unsigned long long*>(d_tmp + 2 * (size_t)n);
  const long long NN = (long long)n * n;
  const int ne = (n + 1) & ~1;
  double floor2 = 0.0;
  dsg_identity<<<(unsigned)((NN + 255) / 256), 256>>>(d_V, n);
  DSG_CK(cudaGetLastError());
  {
    sig2.assign(n, 0.0);
    dsg_column_stats_rect<<<n, DSG_THREADS>>>(d_G, rows, ldg, d_G, d_tmp, d_tmp + n);
    DSG_CK(cudaGetLastError());
    DSG_CK(cudaMemcpy(sig2.data(), d_tmp, n * sizeof(double), cudaMemcpyDeviceToHost));
    double mx = 0;
    for (int j = 0; j < n; j++) if (sig2[j] > mx) mx = sig2[j];
    floor2 = mx * 1e-28;                                   // (1e-14 of the largest column)^2
  }
  for (int sweeps = 0; sweeps < 40; sweeps++) {
    DSG_CK(cudaMemset(d_off, 0, sizeof(unsigned long long)));
    for (int step = 0; step < ne - 1; step++) dsg_jacobi_step_rect<<<ne / 2, DSG_THREADS>>>(d_G, rows, ldg, d_V, n, step, d_off, floor2);
    DSG_CK(cudaGetLastError());
    unsigned long long bits = 0;
    DSG_CK(cudaMemcpy(&bits, d_off, sizeof bits, cudaMemcpyDeviceToHost));
    double off;
    memcpy(&off, &bits, sizeof off);
    if (sweeps_total) (*sweeps_total)++;
    if (off < 1e-15) break;
  }
  sig2.assign(n, 0.0); gb.assign(n, 0.0);
  if (d_rhs) {
    dsg_column_stats_rect<<<n, DSG_THREADS>>>(d_G, rows, ldg, d_rhs, d_tmp, d_tmp + n);
  } else {
    dsg_column_stats_rect<<<n, DSG_THREADS>>>(d_G, rows, ldg, d_G, d_tmp, d_tmp + n);      // gb unused
  }
  DSG_CK(cudaGetLastError());
  DSG_CK(cudaMemcpy(sig2.data(), d_tmp, n * sizeof(double), cudaMemcpyDeviceToHost));
  DSG_CK(cudaMemcpy(gb.data(), d_tmp + n, n * sizeof(double), cudaMemcpyDeviceToHost));
done:
  return rc;
}

// kind 0: analysis Nyquist(M) design; kind 1: synthesis Nyquist(M) design from h_in.  proto_out [L]; *path_out = 3 or 4.
cudaError_t design_prototype_nyquist(int kind, const double* h_in, int M, int m, int r, double wp_factor, int tau, double tolerance,
                                     double* proto_out, int* path_out, int* sweeps_out) {
  const int L = M * m, D = M >> r;
  if (tau < 0) tau = L / 2;
  const int k = kind == 0 ? m : 2 * m;                       // constraint columns
  const long long LL = (long long)L * L;
  cudaError_t rc = cudaSuccess;
  int sweeps = 0, path = 0;
  // host side: the (small) constraint matrix Hc [L][k] and c0
  std::vector<double> Hc((size_t)L * k, 0.0), c0(k, 0.0);
  if (kind == 0) {
    for (int n = 0; n < m; n++) Hc[(size_t)(n * M) * k + n] = 1.0;
    c0[m / 2] = 1.0 / M;
  } else {
    for (int n = 0; n < 2 * m; n++) {
      const int lo = (1 + (n - m) * M) > 0 ? 1 + (n - m) * M : 0, hi = n * M < m * M - 1 ? n * M : m * M - 1;
      for (int kk = lo; kk <= hi; kk++) Hc[(size_t)kk * k + n] = h_in[n * M - kk];
    }
    c0[m] = (double)D / M;
  }
  double *d_A = 0, *d_Q = 0, *d_b = 0, *d_scr = 0, *d_G = 0, *d_V = 0, *d_Hc = 0, *d_rhs = 0, *d_tmp = 0, *d_h = 0, *d_rr = 0;
  double *d_B = 0, *d_T = 0, *d_V2 = 0, *d_x = 0, *d_y = 0, *d_z = 0;
  int* d_idx = 0;
  std::vector<double> sig2, gb, coef, xh(L);
  std::vector<int> nullidx;
  DSG_CK(cudaMalloc(&d_A, LL * sizeof(double)));
  DSG_CK(cudaMalloc(&d_Q, LL * sizeof(double)));
  DSG_CK(cudaMalloc(&d_scr, LL * sizeof(double)));
  DSG_CK(cudaMalloc(&d_b, L * sizeof(double)));
  DSG_CK(cudaMalloc(&d_G, (size_t)(L + k) * L * sizeof(double)));
  DSG_CK(cudaMalloc(&d_V, LL * sizeof(double)));
  DSG_CK(cudaMalloc(&d_Hc, (size_t)L * k * sizeof(double)));
  DSG_CK(cudaMalloc(&d_rhs, (size_t)(L + k) * sizeof(double)));
  DSG_CK(cudaMalloc(&d_tmp, (3 * (size_t)L + 16) * sizeof(double)));
  DSG_CK(cudaMalloc(&d_h, L * sizeof(double)));
  DSG_CK(cudaMalloc(&d_rr, L * sizeof(double)));
  DSG_CK(cudaMalloc(&d_B, LL * sizeof(double)));
  DSG_CK(cudaMalloc(&d_T, LL * sizeof(double)));
  DSG_CK(cudaMalloc(&d_V2, LL * sizeof(double)));
  DSG_CK(cudaMalloc(&d_x, L * sizeof(double)));
  DSG_CK(cudaMalloc(&d_y, L * sizeof(double)));
  DSG_CK(cudaMalloc(&d_z, L * sizeof(double)));
  DSG_CK(cudaMalloc(&d_idx, L * sizeof(int)));
  // passband matrices A, b (PrototypeDesignBase::_calculateAb, :223-246) and the in-band aliasing matrix C (:248-272)
  dsg_analysis_matrices<<<(unsigned)((LL + 255) / 256), 256>>>(d_scr, d_A, d_Q, d_b, L, D, M_PI / (wp_factor * M), tau);
  DSG_CK(cudaGetLastError());
  if (kind == 1) {
    // residual aliasing matrix P (:836-871) replaces C as the quadratic; E, f are not used by this design
    DSG_CK(cudaMemcpy(d_h, h_in, L * sizeof(double), cudaMemcpyHostToDevice));
    dsg_autocorr<<<(L + 127) / 128, 128>>>(d_h, d_rr, L);
    dsg_synthesis_matrices<<<(unsigned)((LL + 255) / 256), 256>>>(d_scr, d_T, d_Q, d_y, d_h, d_rr, L, M, m, D, 0.0, tau);
    DSG_CK(cudaGetLastError());
  }
  DSG_CK(cudaMemcpy(d_Hc, Hc.data(), Hc.size() * sizeof(double), cudaMemcpyHostToDevice));
  {
    // ---- K^T = [Hc^T; Q] and its decomposition: condition number (:579-609), and -- on the singular path -- x_pt and B
    std::vector<double> dp((size_t)L + k, 0.0);
    for (int i = 0; i < k; i++) dp[i] = c0[i];
    DSG_CK(cudaMemcpy(d_rhs, dp.data(), dp.size() * sizeof(double), cudaMemcpyHostToDevice));
    dsg_stack<<<dim3((L + k + 127) / 128, L), 128>>>(d_G, d_Hc, d_Q, L, k);
    DSG_CK(cudaGetLastError());
    DSG_CK(dsg_svd_rect(d_G, L + k, L + k, d_V, L, d_rhs, d_tmp, sig2, gb, &sweeps));
    double s2max = 0, s2min = sig2[0];
    for (int j = 0; j < L; j++) { if (sig2[j] > s2max) s2max = sig2[j]; if (sig2[j] < s2min) s2min = sig2[j]; }
    const double cond = s2min > 0 ? sqrt(s2max / s2min) : HUGE_VAL;
    path = cond < 1.0 / tolerance ? 4 : 3;
    if (getenv("BTK_DESIGN_DEBUG")) fprintf(stderr, "nyquist design kind %d: cond %.6g -> path %d (sweeps %d)\n", kind, cond, path, sweeps);
    if (path == 4) {
      // ---- x_pt = pinv(Hc^T) c0 and the null space of Hc^T: decomposition of Hc^T [k x L]
      DSG_CK(cudaMemcpy(d_rhs, c0.data(), k * sizeof(double), cudaMemcpyHostToDevice));
      // Hc is [L][k] row-major = Hc^T [k x L] column-major with leading dimension k
      DSG_CK(cudaMemcpy(d_G, d_Hc, (size_t)L * k * sizeof(double), cudaMemcpyDeviceToDevice));
      DSG_CK(dsg_svd_rect(d_G, k, k, d_V, L, d_rhs, d_tmp, sig2, gb, &sweeps));
      s2max = 0;
      for (int j = 0; j < L; j++) if (sig2[j] > s2max) s2max = sig2[j];
    }
    const double smax = sqrt(s2max);
    coef.assign(L, 0.0);
    nullidx.clear();
    for (int j = 0; j < L; j++) {
      const double sv = sqrt(sig2[j]);
      if (smax > 0 && sv / smax >= tolerance && sv > 0) coef[j] = gb[j] / sig2[j];
      else nullidx.push_back(j);
    }
    DSG_CK(cudaMemcpy(d_tmp + 2 * (size_t)L, coef.data(), L * sizeof(double), cudaMemcpyHostToDevice));
    dsg_combine<<<(L + 127) / 128, 128>>>(d_V, d_tmp + 2 * (size_t)L, d_x, L);               // x_pt
    DSG_CK(cudaGetLastError());
  }
  {
    const int n0 = (int)nullidx.size();
    if (n0 > 0) {
      const double* d_Mat = path == 4 ? d_Q : d_A;                 // the quadratic minimised inside the null space
      DSG_CK(cudaMemcpy(d_idx, nullidx.data(), n0 * sizeof(int), cudaMemcpyHostToDevice));
      dsg_gather_cols<<<dim3((L + 127) / 128, n0), 128>>>(d_B, d_V, d_idx, L, n0);             // B [L x n0]
      // T = Mat B [L x n0];  Mt = B^T T [n0 x n0] (into d_G);  y = Mat x_pt;  rhs = B^T (Mat x_pt)  or  B^T (b - Mat x_pt)
      dsg_gemm<<<dim3((L + 127) / 128, n0), 128>>>(d_T, L, d_Mat, 1, L, d_B, 1, L, L, n0, L);
      dsg_gemm<<<dim3((n0 + 127) / 128, n0), 128>>>(d_G, n0, d_B, L, 1, d_T, 1, L, n0, n0, L);
      dsg_gemm<<<dim3((L + 127) / 128, 1), 128>>>(d_y, L, d_Mat, 1, L, d_x, 1, L, L, 1, L);
      DSG_CK(cudaGetLastError());
      if (path == 3) {
        std::vector<double> yv(L), bv(L);
        DSG_CK(cudaMemcpy(yv.data(), d_y, L * sizeof(double), cudaMemcpyDeviceToHost));
        DSG_CK(cudaMemcpy(bv.data(), d_b, L * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < L; i++) yv[i] = bv[i] - yv[i];
        DSG_CK(cudaMemcpy(d_y, yv.data(), L * sizeof(double), cudaMemcpyHostToDevice));
      }
      dsg_gemm<<<dim3((n0 + 127) / 128, 1), 128>>>(d_rhs, n0, d_B, L, 1, d_y, 1, L, n0, 1, L);
      DSG_CK(cudaGetLastError());
      DSG_CK(dsg_svd_rect(d_G, n0, n0, d_V2, n0, d_rhs, d_tmp, sig2, gb, &sweeps));
      double s2max = 0;
      for (int j = 0; j < n0; j++) if (sig2[j] > s2max) s2max = sig2[j];
      if (getenv("BTK_DESIGN_DEBUG")) {
        double s2min = s2max; int nsmall = 0;
        for (int j = 0; j < n0; j++) { if (sig2[j] < s2min) s2min = sig2[j]; if (sqrt(sig2[j] / s2max) <= tolerance) nsmall++; }
        fprintf(stderr, "  null space %d, reduced matrix: smax %.6g smin %.6g, %d below the tolerance (sweeps %d)\n", n0, sqrt(s2max), sqrt(s2min), nsmall, sweeps);
      }
      const double smax = sqrt(s2max);
      // The singular path leaves the components below the tolerance UNSCALED (:556-558): x += v_j (u_j^T rhs).  The reduced
      // matrix B^T A B is symmetric positive semi-definite, so u_j = v_j; v_j comes out of the rotations orthonormal
      // whatever s_j is, whereas g_j / s_j is rounding noise for the vanishing singular values (and the reference's GSL
      // delivers an orthonormal U there: the component of a rhs in the range is then zero, as with v_j).
      std::vector<double> vr(n0, 0.0);
      if (path == 3) {
        dsg_gemm<<<dim3((n0 + 127) / 128, 1), 128>>>(d_z, n0, d_V2, n0, 1, d_rhs, 1, n0, n0, 1, n0);       // V^T rhs
        DSG_CK(cudaGetLastError());
        DSG_CK(cudaMemcpy(vr.data(), d_z, n0 * sizeof(double), cudaMemcpyDeviceToHost));
      }
      coef.assign(n0, 0.0);
      for (int j = 0; j < n0; j++) {
        const double sv = sqrt(sig2[j]);
        if (smax > 0 && sv / smax > tolerance && sv > 0) coef[j] = gb[j] / sig2[j];           // V S^-1 U^T rhs
        else if (path == 3) coef[j] = vr[j];
      }
      DSG_CK(cudaMemcpy(d_tmp + 2 * (size_t)L, coef.data(), n0 * sizeof(double), cudaMemcpyHostToDevice));
      dsg_combine<<<(n0 + 127) / 128, 128>>>(d_V2, d_tmp + 2 * (size_t)L, d_z, n0);
      dsg_gemm<<<dim3((L + 127) / 128, 1), 128>>>(d_y, L, d_B, 1, L, d_z, 1, n0, L, 1, n0);      // B z
      DSG_CK(cudaGetLastError());
      std::vector<double> bz(L);
      DSG_CK(cudaMemcpy(xh.data(), d_x, L * sizeof(double), cudaMemcpyDeviceToHost));
      DSG_CK(cudaMemcpy(bz.data(), d_y, L * sizeof(double), cudaMemcpyDeviceToHost));
      for (int i = 0; i < L; i++) proto_out[i] = path == 4 ? xh[i] - bz[i] : xh[i] + bz[i];
    } else {
      DSG_CK(cudaMemcpy(proto_out, d_x, L * sizeof(double), cudaMemcpyDeviceToHost));
    }
  }
done:
  if (path_out) *path_out = path;
  if (sweeps_out) *sweeps_out = sweeps;
  cudaFree(d_A); cudaFree(d_Q); cudaFree(d_scr); cudaFree(d_b); cudaFree(d_G); cudaFree(d_V); cudaFree(d_Hc); cudaFree(d_rhs);
  cudaFree(d_tmp); cudaFree(d_h); cudaFree(d_rr); cudaFree(d_B); cudaFree(d_T); cudaFree(d_V2); cudaFree(d_x); cudaFree(d_y);
  cudaFree(d_z); cudaFree(d_idx);
  return rc;
}

}  // namespace btk
