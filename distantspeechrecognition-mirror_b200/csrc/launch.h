// launch.h -- host-callable launchers of the sm_100a kernels (defined in kern_m*.cu / kern_misc.cu).
#pragma once

#include <cuda_runtime.h>

#include "staged_tiles.cuh"

namespace btk {

// Returns cudaSuccess or the launch error; cudaErrorInvalidValue when (M, R) is not compiled in.
cudaError_t launch_chain(int M, int R, const ChainParams& p, int n_work, cudaStream_t st);
cudaError_t launch_analysis(int M, int R, const AnalysisParams& p, int n_work, cudaStream_t st);
cudaError_t launch_synthesis(int M, int R, const SynthesisParams& p, int n_work, cudaStream_t st);
// The warp-specialised producer / consumer form of the fused chain (chain_ws.cuh).  cudaErrorInvalidValue when the shape
// has no such kernel (its stages do not fit shared memory): the caller then uses launch_chain.  p.cluster CTAs share a
// work item (channel split); n_work counts work items, not CTAs.
cudaError_t launch_chain_ws(int M, int R, const ChainParams& p, int n_work, cudaStream_t st);
int chain_ws_frames_per_iter(int M, int R, int m);       // <0: no warp-specialised kernel for this shape
bool chain_ws_cluster_ok(int M, int R, int m, int S);    // may S CTAs share a work item?
bool fb_supported(int M, int R);
// dynamic shared memory the filter-bank kernels need for (M, R, m); <0 when unsupported
int fb_smem_bytes(int M, int R, int m);
int fb_frames_per_iter(int M, int R);
int chain_frames_per_iter(int M, int R, int m);

// Y[f][s] = sum_c conj(w[s][c]) X[f][s][c]        (beamformer.cc:1181-1194)
cudaError_t launch_beamform(const cf* snap, const cf* w, cf* Y, long long F, int B, int C, cudaStream_t st);
// R[s] += sum_f wt[f] x x^H (conj) or x x^T     (beamformer.cc:142-163 / subbandBeamforming.py:1170-1175)
cudaError_t launch_covariance(const cf* snap, const double* wt, double2* Rout, long long F, int B, int C, int conj,
                              cudaStream_t st);
// the same contraction on tcgen05 tensor cores (TF32 hi/lo split inputs, FP32 accumulation in TMEM): kern_cov_tc.cu
cudaError_t launch_covariance_tc(const cf* snap, const double* wt, double2* Rout, long long F, int B, int C, int conj,
                                 cudaStream_t st);
// batched form: recording i of n reads its snapshots at snap + recs[i].snap_off, its frame weights at wt + recs[i].wt_off,
// recs[i].F frames, and writes Rout + i B C C (device array of descriptors; one launch, grid.z = recording)
struct CovRec { long long snap_off, wt_off, F; };
cudaError_t launch_covariance_tc_batch(const cf* snap, const double* wt, double2* Rout, const CovRec* recs, int n, long long Fmax,
                                       int B, int C, int conj, cudaStream_t st);
// per-bin MVDR solve (beamformer.cc:2392-2446); Rn [n][B][C][C], d [B][C] (one manifold for all n) -> w [n][B][C];
// fallback[n][B] flags.  One CTA per (recording, bin): n = 1 is the single-recording call.
cudaError_t launch_mvdr_solve(const double2* Rn, const double2* d, double2* w, int* fallback, int B, int C,
                              double dThreshold, cudaStream_t st, int n = 1);
// Hermitian positive (semi)definite matrices only (the adaptive batch): L D L^H on the packed lower triangle, see kern_misc.cu
cudaError_t launch_mvdr_chol(const double2* Rn, const double2* d, double2* w, int* fallback, int B, int C, double dThreshold,
                             cudaStream_t st, int n);

// Beamformer output + Zelinski post-filter on stored snapshots (postfilter/postfilter.cc:30-222, 428-500): Y [F][B]
// post-filtered in place semantics (Y is written, then scaled), Wout [F][B] or NULL; stat = scratch of
// zelinski_scratch_bytes(F, B) bytes ([F][B] float4 statistics + one double4 per bin and segment of at least 32 frames).
inline size_t zelinski_scratch_bytes(long long F, int B) { return (size_t)(F * B + 2) * 16 + (size_t)((F + 31) / 32 + 1) * B * 32; }
cudaError_t launch_beamform_zelinski(const cf* snap, const cf* w, const cf* ta, cf* Y, float4* stat, float* Wout, long long F,
                                     int B, int C, double alpha, int type, int min_frames, cudaStream_t st);

// de Haan prototype design in fp64 on the device (kern_design.cu): kind 0 = analysis h = pinv(A + C) b, kind 1 = synthesis
// g = pinv(E + v P) f from h_in (modulated/prototypeDesign.cc:223-272, 640-712, 836-901).  Host pointers.
cudaError_t design_prototype(int kind, const double* h_in, int M, int m, int r, double v, double wp_factor, int tau, double tolerance,
                             double* proto_out, double* err, int* sweeps_out);

// Nyquist(M)-constrained designs (prototypeDesign.cc:955-1119): kind 0 analysis, 1 synthesis from h_in; *path = 3 | 4
cudaError_t design_prototype_nyquist(int kind, const double* h_in, int M, int m, int r, double wp_factor, int tau, double tolerance,
                                     double* proto_out, int* path_out, int* sweeps_out);

// device-resident MVDR adaptation helpers (kern_misc.cu): diagonal loading, chain weight table from device weights
cudaError_t launch_diag_load(double2* Rn, int B, int C, float load_abs, double load_rel, cudaStream_t st);
cudaError_t launch_weight_table(const double2* w, const int* binmap, cf* gam, int M, int C, int Cpad, cudaStream_t st, int n = 1);

// Raw PCM -> float32, element for element (bit-exact integer -> float):
//   fmt 1: int16 little endian (what sf_readf_float returns with SFC_SET_NORM_FLOAT off, feature/feature.cc:273, 868-896)
//   fmt 2: packed 24-bit big endian, sign extended (Conversion24bit2Float::next, feature/feature.cc:190-217;
//          the Mark-III / Mark-IV frame payload, driver/mk4_common.h:50-54)
// src must be 8-byte (fmt 1) / 4-byte (fmt 2) aligned, dst 16-byte aligned.
cudaError_t launch_ingest(int fmt, const void* src, float* dst, long long n, cudaStream_t st);

}  // namespace btk
