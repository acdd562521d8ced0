/* btkb200.h -- C ABI of the B200-native subband front end (analysis bank -> SubbandDS / SubbandMVDR ->
 * synthesis bank).  Plain pointers and sizes only; no C++/torch types; no exceptions cross this boundary
 * (status codes + btkb200_last_error).  The drop-in C++ stream classes (host/btk_streams.h), the Python
 * module and bench.py all sit on top of exactly these entry points.
 *
 * The reference has no FFI of its own: its boundary is the C++ virtual-class contract of btk/stream,
 * btk/modulated and btk/beamformer plus the SWIG projection of it.  Each entry point below
 * cites the reference interface it replaces (paths relative to /root/reference/btk).
 *
 * Layouts
 *   pcm      interleaved float32 [T][C]          (IterativeSampleFeature's buffer, feature/feature.cc:868-896;
 *                                                  values un-normalised, int16 range, feature.cc:273)
 *   snap     complex64 [F][B][C], B = M/2+1      (SnapShotArray::update layout, beamformer/beamformer.cc:82-90;
 *                                                  bins above M/2 are the conjugate mirror and are not stored)
 *   Y        complex64 [F][B]                    (SubbandDS/SubbandMVDR::next output, beamformer.cc:1181-1194)
 *   out      float32 [nblk*D]                    (OverSampledDFTSynthesisBank::next output, concatenated)
 *   weights  complex128 [B][C] as (re, im) pairs (beamformerWeights::_wq / SubbandMVDR::_wmvdr)
 *   R        complex128 [C][C] row-major per bin (SubbandMVDR::_R[fbinX])
 */
#ifndef BTKB200_H
#define BTKB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BTKB200_OK 0
#define BTKB200_EINVAL 1       /* bad argument / size mismatch  -> jdimension_error / jconsistency_error upstream */
#define BTKB200_ESTATE 2       /* call-order error              -> j_error upstream (beamformer.cc:1140-1143,2587-2594) */
#define BTKB200_ECUDA 3        /* CUDA runtime failure; message in btkb200_last_error */
#define BTKB200_ENOMEM 4
#define BTKB200_EUNSUPPORTED 5 /* (M, r) outside the compiled kernel set */

typedef struct btkb200_plan btkb200_plan;

typedef struct btkb200_info {
  unsigned M, m, r, R, D, N, B, C, dct;
  unsigned pd_analysis, pd_synthesis, laN; /* modulated/modulated.cc:278-296 */
  int device;
  int has_weights; /* 0 none, 1 delay-and-sum, 2 user / MVDR */
} btkb200_info;

/* Number of CUDA devices visible; <0 on error. */
int btkb200_device_count(void);
const char* btkb200_version(void);
/* Message of the last failure on this plan (or of the last failed plan_create when plan == NULL). */
const char* btkb200_last_error(const btkb200_plan* plan);

/* OverSampledDFTFilterBank ctor pair (modulated/modulated.cc:262-300, 359-390, 521-569): one plan holds the
 * analysis prototype h[N], the synthesis prototype g[N] (either may be NULL if that side is not used), the
 * geometry (M, m, r, delayCompensationType) and the channel count C of the beamformer
 * (SubbandBeamformer::setChannel, beamformer.cc:1017-1020).  gain = synthesis gainFactor (modulated.h:331-338).
 * Prototypes are copied (modulated.cc:275-276). */
int btkb200_plan_create(btkb200_plan** plan, unsigned M, unsigned m, unsigned r, unsigned dct, unsigned C,
                        const double* h, const double* g, int gain, int device);
void btkb200_plan_destroy(btkb200_plan* plan);
int btkb200_plan_info(const btkb200_plan* plan, btkb200_info* info);
/* ceil(T/D): SampleFeature::next block rule (feature/feature.cc:627-641). */
long btkb200_nblk(const btkb200_plan* plan, long T);
/* nblk + pd - laN frames (modulated.cc:461-516). */
long btkb200_analysis_frames(const btkb200_plan* plan, long T);
/* F - pd_s (modulated.cc:626-642). */
long btkb200_synthesis_frames(const btkb200_plan* plan, long F);
/* Output frames of the whole chain = synthesis_frames(analysis_frames(T)): nblk(T), or nblk(T)+1 for
 * delayCompensationType 2 with an odd m*R (modulated.cc:278-296: pd - laN - pd_synthesis = 1 there). */
long btkb200_chain_frames(const btkb200_plan* plan, long T);

/* ---- weights ------------------------------------------------------------------------------------------ */
/* SubbandDS::calcArrayManifoldVectors -> beamformerWeights::calcMainlobe, halfBandShift=false
 * (beamformer.cc:1087-1091, 531-594).  EINVAL if n_delays != C (jdimension_error, :533-535). */
int btkb200_set_ds_weights(btkb200_plan* plan, double sample_rate, const double* delays, unsigned n_delays);
/* SubbandDS::calcArrayManifoldVectors2 / calcArrayManifoldVectorsN -> beamformerWeights::calcMainlobe2 / calcMainlobeN,
 * halfBandShift=false (beamformer.cc:1100-1121, 603-735, calcNullBeamformer :315-397): unit gain towards delaysT[C], nulls
 * towards the NC-1 interferers delaysJ[NC-1][C].  EINVAL if n_delays != C or NC outside [2, C] (jdimension_error,
 * :634-640).  For NC > 2 the reference inverts C^H C with its single-precision SVD; here in double (differs by the
 * float SVD's rounding). */
int btkb200_set_null_weights(btkb200_plan* plan, double sample_rate, const double* delaysT, unsigned n_delays,
                             const double* delaysJ, unsigned NC);
/* Far-field delays of the shipped drivers, seconds (src/superdirectiveBeamformer.cc:118-137 calcDelaysPolar2; micpos
 * [n][3] in mm, the direction cosines and the quotient in single precision as there). */
int btkb200_calc_delays_polar(float azimuth, float elevation, const double* micpos, unsigned n, double* delays);
/* calcAllDelays (beamformer.cc:1214-1231): |micpos| / c minus the middle element's; the source position is ignored, as in
 * the reference. */
int btkb200_calc_all_delays(double x, double y, double z, const double* micpos, unsigned n, double* delays);
/* Install arbitrary per-bin weights [B][C] (re,im): what SubbandMVDR::next applies (beamformer.cc:2616-2630). */
int btkb200_set_weights(btkb200_plan* plan, const double* w);
/* Weights currently applied by beamform/chain ([B][C] re,im): SubbandDS::getWeights / SubbandMVDR::getMVDRWeights. */
int btkb200_get_weights(const btkb200_plan* plan, double* w);
/* The delay-and-sum manifold wq ([B][C]) kept beside MVDR weights (beamformerWeights::wq_f). */
int btkb200_get_manifold(const btkb200_plan* plan, double* w);
/* The array manifold ([B][C]): the target's delay-and-sum vector that beamformerWeights keeps as _ta (beamformer.cc:583,
 * 992-997) and the post-filter time-aligns with; differs from the quiescent weights only after btkb200_set_null_weights. */
int btkb200_get_array_manifold(const btkb200_plan* plan, double* w);

/* ---- MVDR --------------------------------------------------------------------------------------------- */
/* SubbandMVDR::setNoiseSpatialSpectralMatrix (beamformer.cc:2454-2477); EINVAL on a shape mismatch (the
 * reference returns false). */
int btkb200_set_covariance(btkb200_plan* plan, unsigned bin, const double* R, unsigned rows, unsigned cols);
int btkb200_get_covariance(const btkb200_plan* plan, unsigned bin, double* R);
/* SubbandMVDR::setDiffuseNoiseModel (beamformer.cc:2486-2553); micpos [n_mics][3] in mm. */
int btkb200_set_diffuse_noise_model(btkb200_plan* plan, const double* micpos, unsigned n_mics, double sample_rate,
                                    double sspeed);
/* SubbandMVDR::setAllLevelsOfDiagonalLoading / setLevelOfDiagonalLoading (beamformer.cc:2555-2581). */
int btkb200_diag_load(btkb200_plan* plan, float w);
int btkb200_diag_load_bin(btkb200_plan* plan, unsigned bin, float w);
/* SubbandMVDR::divideAllNonDiagonalElements (beamformer.h:362-378). */
int btkb200_divide_nondiagonal(btkb200_plan* plan, float mu);
/* SubbandMVDR::calcMVDRWeights (beamformer.cc:2392-2446): w[0] = ones; for s >= 1
 * w_s = Rinv^H d / (d^H Rinv d) (d = wq[s]).  The per-bin systems are solved on the device in double
 * precision (pivoted LU) instead of the reference's single-precision SVD pseudoinverse (:253-305); a bin
 * whose smallest pivot magnitude is below dThreshold falls back to the identity like the reference does
 * when pseudoinverse() reports failure (:2425-2427).  *n_fallback (optional) = bins that fell back.
 * ESTATE if no covariance (jallocation_error, :2394-2397) or no manifold (j_error, :2398-2401). */
int btkb200_solve_mvdr(btkb200_plan* plan, double sample_rate, double dThreshold, int* n_fallback);

/* ---- staged path, HOST buffers (copies inside) ---------------------------------------------------------- */
/* C x OverSampledDFTAnalysisBank::next over a whole recording + SnapShotArray::update
 * (modulated.cc:412-516, beamformer.cc:82-90).  snap must hold analysis_frames(T)*B*C complex64. */
int btkb200_analysis(btkb200_plan* plan, const float* pcm, long T, float* snap, long* n_frames);
/* SubbandDS::next / SubbandMVDR::next zdotc loop (beamformer.cc:1181-1194, 2616-2630). */
int btkb200_beamform(btkb200_plan* plan, const float* snap, long F, float* Y);
/* OverSampledDFTSynthesisBank::next over a whole stream (modulated.cc:595-664). */
int btkb200_synthesis(btkb200_plan* plan, const float* Y, long F, float* out, long* n_out_frames);
/* Weighted Gram matrices R[s] = sum_f wt[f] x_f x_f^H (conjugate != 0; lib/subbandBeamforming.py:1170-1175)
 * or x_f x_f^T (conjugate == 0; SpectralMatrixArray::update, beamformer.cc:142-163) for bins 0..M/2.
 * R: [B][C][C] complex128.  EINVAL for a negative or non-finite frame weight. */
int btkb200_covariance(btkb200_plan* plan, const float* snap, long F, const double* frame_weights, int conjugate,
                       double* R);

/* Adaptive (sample-covariance) MVDR in one call, entirely on the device: analysis of pcm [T][C] -> per-bin
 * spatial covariance with exponential forgetting -> the plan's noise matrices (as if setNoiseSpatialSpectralMatrix had
 * been called for every bin).  Follows SubbandBeamformerMVDR.updateSx (lib/subbandBeamforming.py:1138, 1170-1175):
 * S = x x^H at frame 0, then S <- ff S + (1-ff) x x^H while frame <= last_frame (last_frame < 0: all frames);
 * conjugate == 0 gives the C++ SpectralMatrixArray::update flavour x x^T with R <- mu R + (1-mu) x x^T from R = 0
 * (beamformer.cc:142-163).  Follow with btkb200_diag_load and btkb200_solve_mvdr; no subband data touches the host. */
int btkb200_estimate_covariance(btkb200_plan* plan, const float* pcm, long T, double forget, long last_frame,
                                int conjugate);

/* ---- prototype design ------------------------------------------------------------------------------ */
/* AnalysisOversampledDFTDesign(M, m, r, wpFactor, tau_h).design(tolerance) (modulated/prototypeDesign.cc:223-272, 611-712):
 * h = pinv(A + C) b, singular values below tolerance * s_max dropped; tau < 0 = M m / 2.  h: M*m doubles; err (or NULL)
 * receives calcError(): passband response error, in-band aliasing distortion, their sum (dB).  fp64 on the device. */
int btkb200_design_analysis_prototype(unsigned M, unsigned m, unsigned r, double wp_factor, int tau, double tolerance,
                                      int device, double* h, double* err);
/* SynthesisOversampledDFTDesign(h, M, m, r, v, wpFactor, tau_g).design(tolerance) (prototypeDesign.cc:768-901):
 * g = pinv(E + v P) f.  err: total response error, residual aliasing distortion, eps_t + v eps_r (dB). */
int btkb200_design_synthesis_prototype(const double* h, unsigned M, unsigned m, unsigned r, double v, double wp_factor,
                                       int tau, double tolerance, int device, double* g, double* err);

/* AnalysisNyquistMDesign(M, m, r, wpFactor, tau_h).design(tolerance) (prototypeDesign.cc:955-1001): the in-band aliasing
 * h^T C h minimised subject to the Nyquist(M) constraint h[n M] = delta(n - m/2) / M ("alternate solution 4", :361-470), or --
 * when cond([F^T; C]) >= 1/tolerance (:579-609) -- the passband error minimised inside the numerical null space of the
 * stacked constraints ("alternate solution 3", :481-577).  *path (or NULL) receives 3 or 4.  fp64 on the device. */
int btkb200_design_analysis_nyquist(unsigned M, unsigned m, unsigned r, double wp_factor, int tau, double tolerance,
                                    int device, double* h, int* path);
/* SynthesisNyquistMDesign(h, M, m, r, wpFactor, tau_g).design(tolerance) (prototypeDesign.cc:1003-1119): the residual
 * aliasing g^T P g minimised subject to the 2 m total-response constraints H^T g = c0 (:1073-1089); same two paths. */
int btkb200_design_synthesis_nyquist(const double* h, unsigned M, unsigned m, unsigned r, double wp_factor, int tau,
                                     double tolerance, int device, double* g, int* path);

/* ---- SubbandGSC with fixed active weights ------------------------------------------------------- */
/* SubbandGSC::calcGSCWeights (beamformer.cc:1373-1377): delay-and-sum quiescent vectors + one blocking matrix per bin
 * (_calcBlockingMatrix, :398-479, NC = 1); the active weights start at zero.  EINVAL for a single channel (:536-539). */
int btkb200_gsc_calc_weights(btkb200_plan* plan, double sample_rate, const double* delays, unsigned n_delays);
/* SubbandGSC::setActiveWeights_f (:1425-1433 -> calcSidelobeCancellerP_f :761-783): packed = (re, im) x (C-1).
 * ESTATE before gsc_calc_weights (j_error), EINVAL for a wrong length (jdimension_error). */
int btkb200_gsc_set_active_weights(btkb200_plan* plan, unsigned bin, const double* packed, unsigned n);
/* SubbandGSC::zeroActiveWeights (:1435-1447). */
int btkb200_gsc_zero_active_weights(btkb200_plan* plan);
/* SubbandGSC::getBlockingMatrix(0, bin): [C][C-1] complex128, row major. */
int btkb200_gsc_get_blocking_matrix(const btkb200_plan* plan, unsigned bin, double* B);
/* Installs the weights SubbandGSC::next applies (:1296-1356, calcOutputOfGSC :1251-1289): w = wq - B wa for bins
 * 1..M/2 (divided by ||w|| C when normalize != 0, normalizeWeight(true)), wq alone for bin 0.  Afterwards every
 * beamform / chain entry point computes the GSC output. */
int btkb200_gsc_apply(btkb200_plan* plan, int normalize);

/* ---- Zelinski post-filter ------------------------------------------------------------------- */
/* ZelinskiPostFilter::next over a whole recording (postfilter/postfilter.cc:428-500 -> ZelinskiFilter :153-222 ->
 * ZelinskiFilter_f :56-140), with the beamformer as set by setBeamformer(): snapshots [F][B][C] -> beamformer output
 * (current weights) multiplied by the post-filter gain, Y [F][B] complex64; W [F][B] float32 receives the gains
 * (getPostFilterWeights() of every frame) when not NULL.  alpha = forgetting factor (default 0.6), type =
 * PostfilterType (1 = TYPE_ZELINSKI1_REAL, 2 = TYPE_ZELINSKI1_ABS, 0 = NO_USE_POST_FILTER: gains only), min_frames as
 * in the constructor.  The time alignment uses the array manifold of set_ds_weights (beamformerWeights::arrayManifold).
 * EINVAL for fewer than two channels (jdimension_error, :62-65). */
int btkb200_beamform_zelinski(btkb200_plan* plan, const float* snap, long F, double alpha, int type, int min_frames,
                              float* Y, float* W);
/* analysis -> beamformer -> Zelinski post-filter -> synthesis on the device, host pcm in, host PCM out
 * (the chain of src/superdirectiveBeamformer.cc:150-205 and src/beamformerDS.cc); out holds chain_frames(T)*D floats. */
int btkb200_chain_zelinski(btkb200_plan* plan, const float* pcm, long T, double alpha, int type, int min_frames,
                           float* out);

/* Same for n independent recordings, pipelined (uploads ahead on a copy stream, downloads behind on another). */
int btkb200_chain_zelinski_batch(btkb200_plan* plan, const float* const* pcm, const long* T, int n, double alpha, int type,
                                 int min_frames, float* const* out);

/* ---- fused path ---------------------------------------------------------------------------------------- */
/* pcm -> out through analysis -> weight apply -> synthesis in ONE kernel; out holds chain_frames(T)*D floats.
 * ESTATE if no weights are installed (j_error, beamformer.cc:1140-1143). */
int btkb200_chain(btkb200_plan* plan, const float* pcm, long T, float* out);
/* n independent recordings (ragged lengths allowed), host buffers. */
int btkb200_chain_batch(btkb200_plan* plan, const float* const* pcm, const long* T, int n, float* const* out);
/* Adaptive MVDR over a batch, entirely on the device (BASELINE configs 2-3: "SubbandMVDR with per-subband spatial covariance
 * + diagonal loading", many utterances).  Per recording, in the reference's own call order:
 *   SpectralMatrixArray::update over frames 0..last_frame (conjugate = 0, beamformer.cc:142-163) or updateSx (conjugate = 1,
 *   lib/subbandBeamforming.py:1170-1175)  ->  setNoiseSpatialSpectralMatrix for every bin (beamformer.cc:2454-2477)
 *   ->  setAllLevelsOfDiagonalLoading(load_abs) (:2555-2568; the weight is kept as a float like there) plus, when
 *   load_rel != 0, load_rel * trace(R_s) / C per bin  ->  calcMVDRWeights(dThreshold) (:2392-2446)  ->  the fused
 *   analysis -> weight apply -> synthesis chain with THAT recording's weights.
 * The analysis of the adapting lead-in, the tensor-core covariance, the loading, the per-bin solve and the chain weight
 * table run per recording on the device; one fused-chain launch then covers the whole batch with a weight table per
 * recording.  No subband data, matrices or weights cross PCIe.  n_fallback[i] (or NULL) receives the number of bins of
 * recording i whose inverse failed (identity fallback, :2425-2427).  The array manifold comes from set_ds_weights. */
typedef struct btkb200_mvdr_adapt {
  double forget;     /* forgetting factor of the covariance recursion (mu / ff) */
  long last_frame;   /* frames 0..last_frame adapt; < 0: every frame */
  int conjugate;     /* 1: x x^H (updateSx), 0: x x^T (SpectralMatrixArray::update) */
  double load_abs;   /* setAllLevelsOfDiagonalLoading */
  double load_rel;   /* additional load_rel * trace(R_s) / C */
  double dThreshold; /* calcMVDRWeights */
} btkb200_mvdr_adapt;
int btkb200_mvdr_chain_batch(btkb200_plan* plan, const float* const* pcm, const long* T, int n,
                             const btkb200_mvdr_adapt* cfg, float* const* out, int* n_fallback);
/* Same with raw interleaved PCM in host memory, converted on the device (element order unchanged, exact):
 *   BTKB200_PCM_F32   float32                           (IterativeSampleFeature buffer, feature/feature.cc:868-896)
 *   BTKB200_PCM_S16   int16 little endian               (16-bit WAV payload: what sf_readf_float delivers with
 *                                                        SFC_SET_NORM_FLOAT off, feature/feature.cc:273)
 *   BTKB200_PCM_S24BE packed 24-bit big endian, signed  (Conversion24bit2Float::next, feature/feature.cc:190-217;
 *                                                        Mark-III/IV data frames: 64 ch x 3 bytes, driver/mk4_common.h:50-54)
 * Halves (or cuts by a quarter) the host->device bytes of the end-to-end path, which is PCIe-bound. */
#define BTKB200_PCM_F32 0
#define BTKB200_PCM_S16 1
#define BTKB200_PCM_S24BE 2
int btkb200_chain_batch_pcm(btkb200_plan* plan, const void* const* pcm, int format, const long* T, int n,
                            float* const* out);
/* The conversion alone: n raw samples (host) -> n floats (host), through the device kernels. */
int btkb200_convert_pcm(btkb200_plan* plan, const void* src, int format, long n, float* dst);
/* Same, recordings spread round-robin over n_plans plans living on different devices (no inter-GPU traffic). */
int btkb200_chain_batch_multi(btkb200_plan* const* plans, int n_plans, const float* const* pcm, const long* T, int n,
                              float* const* out);

/* ---- device-resident variants (pointers are DEVICE pointers on the plan's device; stream is a cudaStream_t
 *      passed as void*, NULL = the legacy default stream).  They return after enqueueing the kernels; the descriptor
 *      upload before a batch launch waits for earlier work on `stream`.  One stream per plan: the plan owns the
 *      descriptor and scratch buffers a launch reads, so two streams must not drive the same plan concurrently.
 *      d_pcm need not be 16-byte aligned (an offset view takes the scalar load path). -------- */
int btkb200_chain_batch_dev(btkb200_plan* plan, const float* d_pcm, const long long* pcm_off, const long long* T,
                            const long long* out_off, int n, float* d_out, void* stream);
int btkb200_analysis_dev(btkb200_plan* plan, const float* d_pcm, long T, float* d_snap, void* stream);
int btkb200_beamform_dev(btkb200_plan* plan, const float* d_snap, long F, float* d_Y, void* stream);
int btkb200_beamform_zelinski_dev(btkb200_plan* plan, const float* d_snap, long F, double alpha, int type, int min_frames,
                                  float* d_Y, float* d_W, void* stream);
int btkb200_synthesis_dev(btkb200_plan* plan, const float* d_Y, long F, float* d_out, void* stream);
/* Launch-geometry knobs of the fused chain (measurement aid: A/B runs and tests; results do not depend on them).
 *   BTKB200_TUNE_CHAIN_WS  1 / 0: use / do not use the warp-specialised producer-consumer kernel (chain_ws.cuh); -1 = automatic
 *   BTKB200_TUNE_CLUSTER   n > 0: n CTAs of a thread-block cluster share a work item (channel split); 0 = automatic
 * An impossible value (shape without such a kernel, n not dividing the channel groups) returns BTKB200_EUNSUPPORTED and
 * leaves the knob unchanged.  No reference counterpart: the reference has no launch geometry. */
#define BTKB200_TUNE_CHAIN_WS 1
#define BTKB200_TUNE_CLUSTER 2
int btkb200_plan_tune(btkb200_plan* plan, int knob, int value);
/* what the last fused-chain launch of this plan used: value of the knob after the automatic choice (-1 before any launch) */
int btkb200_plan_tuning(const btkb200_plan* plan, int knob);
/* Kernels enqueued by this plan so far (for launch accounting in bench.py). */
long btkb200_launch_count(const btkb200_plan* plan);
/* Block until all work enqueued by this plan has finished. */
int btkb200_sync(btkb200_plan* plan);

/* Page-locked host buffers for the host-buffer entry points (optional; pageable memory also works). */
void* btkb200_host_alloc(size_t bytes);
void btkb200_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* BTKB200_H */
