// oracle/ref_driver.cc
//
// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Thin extern "C" harness around the UNMODIFIED reference classes, compiled by
// oracle/Makefile together with the reference's own sources (read in place from
// /root/reference/btk) into oracle/_ref/libbtk_ref.so.  Everything numerical
// below is done by the reference:
//   OverSampledDFTAnalysisBank   btk/modulated/modulated.cc:359-516
//   OverSampledDFTSynthesisBank  btk/modulated/modulated.cc:521-674
//   SubbandDS                    btk/beamformer/beamformer.cc:1057-1212
//   SubbandMVDR                  btk/beamformer/beamformer.cc:2321-2635
//   SpectralMatrixArray          btk/beamformer/beamformer.cc:119-163
//   ZelinskiPostFilter           btk/postfilter/postfilter.cc:340-500
//   Analysis/SynthesisOversampledDFTDesign   btk/modulated/prototypeDesign.cc:223-272, 611-951
// This file only feeds them from memory and copies their per-frame outputs out.
// The in-memory source reproduces SampleFeature::next's block/pad rule
// (btk/feature/feature.cc:610-659) because feature.cc itself needs libsndfile.

#include <stdio.h>
#include <string.h>
#include <vector>
#include <exception>

#include "stream/stream.h"
#include "modulated/modulated.h"
#include "beamformer/beamformer.h"
#include "modulated/prototypeDesign.h"

// postfilter/postfilter.cc is part of the build (ZelinskiPostFilter: btk/postfilter/postfilter.cc:30-222, 340-500).
#include "postfilter/postfilter.h"

namespace {

// In-memory block source: one channel of a (possibly interleaved) float buffer.
// Block/pad semantics follow SampleFeature::next (feature.cc:610-659) with
// blockLen == shiftLen == D and padZeros == true.
class MemorySampleFeature : public VectorFloatFeatureStream {
 public:
  MemorySampleFeature(const float* base, long T, unsigned stride, unsigned D)
      : VectorFloatFeatureStream(D, "MemorySampleFeature"), _base(base), _T(T), _stride(stride), _cur(0) {}

  virtual const gsl_vector_float* next(int frameX = -5) {
    if (_endOfSamples) throw jiterator_error("end of samples!");
    if (frameX == _frameX) return _vector;
    if (frameX >= 0 && frameX - 1 != _frameX)
      throw jindex_error("Problem in Feature %s: %d != %d\n", name().c_str(), frameX - 1, _frameX);
    if (_cur >= _T) { _endOfSamples = true; throw jiterator_error("end of samples!"); }
    const long n = (long)size();
    if (_cur + n >= _T) {
      gsl_vector_float_set_zero(_vector);
      for (long i = 0; i < _T - _cur; i++) gsl_vector_float_set(_vector, i, _base[(_cur + i) * _stride]);
    } else {
      for (long i = 0; i < n; i++) gsl_vector_float_set(_vector, i, _base[(_cur + i) * _stride]);
    }
    _cur += n;
    _increment();
    return _vector;
  }
  virtual void reset() { _cur = 0; VectorFloatFeatureStream::reset(); }

 private:
  const float* _base;
  const long _T;
  const unsigned _stride;
  long _cur;
};

// Complex stream fed from a memory buffer of full-M spectra (for synthesis-only runs).
class MemoryComplexFeature : public VectorComplexFeatureStream {
 public:
  MemoryComplexFeature(const double* Y, long F, unsigned M)
      : VectorComplexFeatureStream(M, "MemoryComplexFeature"), _Y(Y), _F(F) {}
  virtual const gsl_vector_complex* next(int frameX = -5) {
    if (frameX == _frameX) return _vector;
    long f = (long)_frameX + 1;
    if (f >= _F) { _endOfSamples = true; throw jiterator_error("end of samples!"); }
    for (unsigned s = 0; s < size(); s++)
      gsl_vector_complex_set(_vector, s, gsl_complex_rect(_Y[2 * (f * size() + s)], _Y[2 * (f * size() + s) + 1]));
    _increment();
    return _vector;
  }

 private:
  const double* _Y;
  const long _F;
};

// Pass-through node that records what the beamformer emitted (and its snapshots).
class RecordingTee : public VectorComplexFeatureStream {
 public:
  RecordingTee(SubbandDS* bf, unsigned M, unsigned C, double* Y, double* snap, long cap)
      : VectorComplexFeatureStream(M, "RecordingTee"), _bf(bf), _C(C), _Y(Y), _snap(snap), _cap(cap), _n(0) {}
  virtual const gsl_vector_complex* next(int frameX = -5) {
    if (frameX == _frameX) return _vector;
    const gsl_vector_complex* y = _bf->next(frameX);
    const unsigned M = size();
    gsl_vector_complex_memcpy(_vector, y);
    if (_n < _cap) {
      if (_Y) for (unsigned s = 0; s < M; s++) {
        gsl_complex z = gsl_vector_complex_get(y, s);
        _Y[2 * (_n * M + s)] = GSL_REAL(z); _Y[2 * (_n * M + s) + 1] = GSL_IMAG(z);
      }
      if (_snap) for (unsigned s = 0; s < M; s++) {
        const gsl_vector_complex* x = _bf->snapShotArray_f(s);
        for (unsigned c = 0; c < _C; c++) {
          gsl_complex z = gsl_vector_complex_get(x, c);
          double* p = _snap + 2 * ((_n * M + s) * _C + c);
          p[0] = GSL_REAL(z); p[1] = GSL_IMAG(z);
        }
      }
    }
    _n++;
    _increment();
    return _vector;
  }
  virtual void reset() { _bf->reset(); VectorComplexFeatureStream::reset(); }
  long frames() const { return _n; }

 private:
  SubbandDS* _bf;
  const unsigned _C;
  double* _Y;
  double* _snap;
  const long _cap;
  long _n;
};

// Oracle "B" (SURVEY 8c): the reference's calcMVDRWeights with the float-SVD
// pseudoinverse (beamformer.cc:253-305) replaced by a double-precision inverse.
// _invR is protected, so a subclass fills it and calls calcMVDRWeights(...,false),
// which is the reference's own "inverse already supplied" path (beamformer.cc:2422).
class MVDRDoubleInverse : public SubbandMVDR {
 public:
  MVDRDoubleInverse(unsigned fftLen) : SubbandMVDR(fftLen, false, "SubbandMVDR") {}
  bool fillDoubleInverses() {
    const unsigned C = chanN();
    bool ok = true;
    for (unsigned s = 1; s <= _fftLen / 2; s++) {
      if (_R[s] == NULL) return false;
      if (_invR[s] == NULL) _invR[s] = gsl_matrix_complex_alloc(C, C);
      ok = invert(_R[s], _invR[s], C) && ok;
    }
    return ok;
  }

 private:
  // Gauss-Jordan with partial pivoting in long double complex arithmetic.
  static bool invert(const gsl_matrix_complex* A, gsl_matrix_complex* inv, unsigned n) {
    typedef long double ld;
    std::vector<ld> ar(n * 2 * n), ai(n * 2 * n);
    for (unsigned i = 0; i < n; i++) for (unsigned j = 0; j < n; j++) {
      gsl_complex z = gsl_matrix_complex_get(A, i, j);
      ar[i * 2 * n + j] = GSL_REAL(z); ai[i * 2 * n + j] = GSL_IMAG(z);
      ar[i * 2 * n + n + j] = (i == j) ? 1.0L : 0.0L; ai[i * 2 * n + n + j] = 0.0L;
    }
    for (unsigned col = 0; col < n; col++) {
      unsigned piv = col; ld best = -1;
      for (unsigned i = col; i < n; i++) {
        ld a = ar[i * 2 * n + col] * ar[i * 2 * n + col] + ai[i * 2 * n + col] * ai[i * 2 * n + col];
        if (a > best) { best = a; piv = i; }
      }
      if (best <= 0) return false;
      if (piv != col) for (unsigned j = 0; j < 2 * n; j++) {
        std::swap(ar[piv * 2 * n + j], ar[col * 2 * n + j]); std::swap(ai[piv * 2 * n + j], ai[col * 2 * n + j]);
      }
      ld pr = ar[col * 2 * n + col], pi = ai[col * 2 * n + col], d = pr * pr + pi * pi;
      ld ir = pr / d, ii = -pi / d;
      for (unsigned j = 0; j < 2 * n; j++) {
        ld xr = ar[col * 2 * n + j], xi = ai[col * 2 * n + j];
        ar[col * 2 * n + j] = xr * ir - xi * ii; ai[col * 2 * n + j] = xr * ii + xi * ir;
      }
      for (unsigned i = 0; i < n; i++) {
        if (i == col) continue;
        ld fr = ar[i * 2 * n + col], fi = ai[i * 2 * n + col];
        if (fr == 0 && fi == 0) continue;
        for (unsigned j = 0; j < 2 * n; j++) {
          ld xr = ar[col * 2 * n + j], xi = ai[col * 2 * n + j];
          ar[i * 2 * n + j] -= fr * xr - fi * xi; ai[i * 2 * n + j] -= fr * xi + fi * xr;
        }
      }
    }
    for (unsigned i = 0; i < n; i++) for (unsigned j = 0; j < n; j++)
      gsl_matrix_complex_set(inv, i, j, gsl_complex_rect((double)ar[i * 2 * n + n + j], (double)ai[i * 2 * n + n + j]));
    return true;
  }
};

// SynthesisOversampledDFTDesign re-declares _singularVals / _scratch / _workSpace (prototypeDesign.h:199-202), hiding the
// base-class vectors, and its constructor never allocates them (prototypeDesign.cc:768-779): _solve() then hands
// uninitialised pointers to the SVD (:886-899).  The members are protected, so a subclass can allocate them; nothing
// else of the reference's design code is touched.
class SynthesisDesignWithWorkspace : public SynthesisOversampledDFTDesign {
 public:
  SynthesisDesignWithWorkspace(const gsl_vector* h, int M, int m, int r, double v, double wp, int tau)
      : SynthesisOversampledDFTDesign(h, M, m, r, v, wp, tau) {
    _singularVals = gsl_vector_calloc(M * m);
    _scratch = gsl_vector_calloc(M * m);
    _workSpace = gsl_vector_calloc(M * m);
  }
  ~SynthesisDesignWithWorkspace() { gsl_vector_free(_singularVals); gsl_vector_free(_scratch); gsl_vector_free(_workSpace); }
};

gsl_vector* make_vector(const double* src, size_t n) {
  gsl_vector* v = gsl_vector_alloc(n);
  for (size_t i = 0; i < n; i++) gsl_vector_set(v, i, src[i]);
  return v;
}

}  // namespace

// Nyquist(M)-constrained designs (prototypeDesign.cc:955-1119): AnalysisNyquistMDesign, then SynthesisNyquistMDesign from its
// h.  The synthesis class inherits the unallocated work vectors of SynthesisOversampledDFTDesign (see above).
class SynthesisNyquistWithWorkspace : public SynthesisNyquistMDesign {
 public:
  SynthesisNyquistWithWorkspace(const gsl_vector* h, int M, int m, int r, double wp, int tau)
      : SynthesisNyquistMDesign(h, M, m, r, wp, tau) {
    _singularVals = gsl_vector_calloc(M * m);
    _scratch = gsl_vector_calloc(M * m);
    _workSpace = gsl_vector_calloc(M * m);
  }
  ~SynthesisNyquistWithWorkspace() { gsl_vector_free(_singularVals); gsl_vector_free(_scratch); gsl_vector_free(_workSpace); }
};

extern "C" int btkref_design_nyquist(int M, int m, int r, double wpFactor, double tolerance, double* h, double* g,
                                     double* err_h, double* err_g) {
  try {
    const int L = M * m;
    AnalysisNyquistMDesign ana(M, m, r, wpFactor, -1);
    const gsl_vector* hv = ana.design(tolerance);
    for (int n = 0; n < L; n++) h[n] = gsl_vector_get(hv, n);
    const gsl_vector* eh = ana.calcError(false);
    if (err_h) for (int k = 0; k < 3; k++) err_h[k] = gsl_vector_get(eh, k);
    if (g) {
      gsl_vector* hc = make_vector(h, L);
      SynthesisNyquistWithWorkspace syn(hc, M, m, r, wpFactor, -1);
      const gsl_vector* gv = syn.design(tolerance);
      for (int n = 0; n < L; n++) g[n] = gsl_vector_get(gv, n);
      const gsl_vector* eg = syn.calcError(false);
      if (err_g) for (int k = 0; k < 2; k++) err_g[k] = gsl_vector_get(eg, k);
      gsl_vector_free(hc);
    }
    return 0;
  } catch (std::exception& e) { fprintf(stderr, "btkref_design_nyquist: %s\n", e.what()); return -1; }
}

extern "C" {

struct btkref_chain_cfg {
  int M, m, r, dct, C;
  double fs;
  int mode;          // 0 = SubbandDS, 1 = SubbandMVDR
  int inverse_kind;  // MVDR: 0 = reference float-SVD pseudoinverse (oracle A), 1 = double inverse (oracle B)
  int noise_model;   // MVDR: 0 = explicit Rn per bin, 1 = setDiffuseNoiseModel(micpos)
  double dThreshold;
  float diag_load;   // applied with setAllLevelsOfDiagonalLoading when != 0
  float divide_mu;   // applied with divideAllNonDiagonalElements when >= 0
  double sspeed;
  int gain;          // synthesis gainFactor
};

// Single-channel analysis.  X: [cap][M][2] doubles.  Returns frames emitted (may exceed cap), <0 on error.
long btkref_analysis(const float* x, long T, const double* h, int M, int m, int r, int dct, double* X, long cap) {
  try {
    const unsigned D = M >> r;
    gsl_vector* proto = make_vector(h, (size_t)M * m);
    VectorFloatFeatureStreamPtr src(new MemorySampleFeature(x, T, 1, D));
    OverSampledDFTAnalysisBankPtr bank(new OverSampledDFTAnalysisBank(src, proto, M, m, r, dct));
    gsl_vector_free(proto);
    long n = 0;
    try {
      for (;;) {
        const gsl_vector_complex* f = bank->next();
        if (n < cap) for (int s = 0; s < M; s++) {
          gsl_complex z = gsl_vector_complex_get(f, s);
          X[2 * (n * M + s)] = GSL_REAL(z); X[2 * (n * M + s) + 1] = GSL_IMAG(z);
        }
        n++;
      }
    } catch (jiterator_error&) {}
    return n;
  } catch (std::exception& e) { fprintf(stderr, "btkref_analysis: %s\n", e.what()); return -1; }
}

// Synthesis of a stored complex stream.  Y: [F][M][2]; out: [cap*D] floats.  Returns frames emitted.
long btkref_synthesis(const double* Y, long F, const double* g, int M, int m, int r, int dct, int gain, float* out, long cap) {
  try {
    const unsigned D = M >> r;
    gsl_vector* proto = make_vector(g, (size_t)M * m);
    VectorComplexFeatureStreamPtr src(new MemoryComplexFeature(Y, F, M));
    OverSampledDFTSynthesisBankPtr bank(new OverSampledDFTSynthesisBank(src, proto, M, m, r, dct, gain));
    gsl_vector_free(proto);
    long n = 0;
    try {
      for (;;) {
        const gsl_vector_float* f = bank->next();
        if (n < cap) for (unsigned d = 0; d < D; d++) out[n * D + d] = gsl_vector_float_get(f, d);
        n++;
      }
    } catch (jiterator_error&) {}
    return n;
  } catch (std::exception& e) { fprintf(stderr, "btkref_synthesis: %s\n", e.what()); return -1; }
}

// Full chain: C analysis banks -> SubbandDS/SubbandMVDR -> (optional) synthesis.
//   pcm   : interleaved [T][C] floats
//   delays: [C] seconds;  Rn: [B][C][C][2] (noise_model 0) ; micpos: [C][3] mm (noise_model 1)
//   snap  : [cap][M][C][2] or NULL;  Y: [cap][M][2] or NULL;  out: [cap_out*D] or NULL (no synthesis when g==NULL)
//   W     : [B][C][2] weights actually used (wq for DS, wmvdr for MVDR) or NULL
// Returns beamformer frames emitted; *n_out = synthesis frames emitted.
long btkref_chain(const btkref_chain_cfg* cfg, const float* pcm, long T, const double* h, const double* g,
                  const double* delays, const double* Rn, const double* micpos,
                  double* snap, double* Y, long cap, float* out, long cap_out, long* n_out, double* W) {
  try {
    const int M = cfg->M, m = cfg->m, r = cfg->r, C = cfg->C;
    const unsigned D = M >> r, B = M / 2 + 1;
    gsl_vector* hp = make_vector(h, (size_t)M * m);
    SubbandDS* bf;
    MVDRDoubleInverse* mv = NULL;
    if (cfg->mode == 1) { mv = new MVDRDoubleInverse(M); bf = mv; } else { bf = new SubbandDS(M, false); }
    for (int c = 0; c < C; c++) {
      VectorFloatFeatureStreamPtr src(new MemorySampleFeature(pcm + c, T, C, D));
      VectorComplexFeatureStreamPtr bank(new OverSampledDFTAnalysisBank(src, hp, M, m, r, cfg->dct));
      bf->setChannel(bank);
    }
    gsl_vector_free(hp);
    gsl_vector* dv = make_vector(delays, C);
    bf->calcArrayManifoldVectors(cfg->fs, dv);
    gsl_vector_free(dv);
    if (mv) {
      if (cfg->noise_model == 1) {
        gsl_matrix* mp = gsl_matrix_alloc(C, 3);
        for (int c = 0; c < C; c++) for (int k = 0; k < 3; k++) gsl_matrix_set(mp, c, k, micpos[c * 3 + k]);
        if (!mv->setDiffuseNoiseModel(mp, cfg->fs, cfg->sspeed)) { gsl_matrix_free(mp); return -2; }
        gsl_matrix_free(mp);
      } else {
        gsl_matrix_complex* Rm = gsl_matrix_complex_alloc(C, C);
        for (unsigned s = 0; s < B; s++) {
          for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) {
            const double* p = Rn + 2 * ((size_t)(s * C + i) * C + j);
            gsl_matrix_complex_set(Rm, i, j, gsl_complex_rect(p[0], p[1]));
          }
          if (!mv->setNoiseSpatialSpectralMatrix(s, Rm)) { gsl_matrix_complex_free(Rm); return -3; }
        }
        gsl_matrix_complex_free(Rm);
      }
      if (cfg->divide_mu >= 0.0f) mv->divideAllNonDiagonalElements(cfg->divide_mu);
      if (cfg->diag_load != 0.0f) mv->setAllLevelsOfDiagonalLoading(cfg->diag_load);
      if (cfg->inverse_kind == 1) {
        if (!mv->fillDoubleInverses()) return -4;
        mv->calcMVDRWeights(cfg->fs, cfg->dThreshold, false);
      } else {
        mv->calcMVDRWeights(cfg->fs, cfg->dThreshold, true);
      }
    }
    if (W) for (unsigned s = 0; s < B; s++) {
      const gsl_vector_complex* w = mv ? mv->getMVDRWeights(s) : bf->getWeights(s);
      for (int c = 0; c < C; c++) {
        gsl_complex z = gsl_vector_complex_get(w, c);
        W[2 * (s * C + c)] = GSL_REAL(z); W[2 * (s * C + c) + 1] = GSL_IMAG(z);
      }
    }
    RecordingTee* tee = new RecordingTee(bf, M, C, Y, snap, cap);
    VectorComplexFeatureStreamPtr teep(tee);
    long nsyn = 0;
    if (g) {
      gsl_vector* gp = make_vector(g, (size_t)M * m);
      OverSampledDFTSynthesisBankPtr syn(new OverSampledDFTSynthesisBank(teep, gp, M, m, r, cfg->dct, cfg->gain));
      gsl_vector_free(gp);
      try {
        for (;;) {
          const gsl_vector_float* f = syn->next();
          if (out && nsyn < cap_out) for (unsigned d = 0; d < D; d++) out[nsyn * D + d] = gsl_vector_float_get(f, d);
          nsyn++;
        }
      } catch (jiterator_error&) {}
    } else {
      try { for (;;) teep->next(); } catch (jiterator_error&) {}
    }
    if (n_out) *n_out = nsyn;
    long nf = tee->frames();
    // bf is owned by nobody's smart pointer (the tee holds a raw pointer): release it here.
    bf->clearChannel();
    delete bf;
    return nf;
  } catch (std::exception& e) { fprintf(stderr, "btkref_chain: %s\n", e.what()); return -1; }
}

// analysis banks -> SubbandDS -> ZelinskiPostFilter -> (optional) synthesis, wired like the shipped drivers
// (src/beamformerDS.cc:150-190, src/superdirectiveBeamformer.cc:150-205).
//   Ypf : [cap][M][2] post-filtered beamformer output or NULL;  Wpf: [cap][M] post-filter gains (real part of wp1) or NULL
// Returns post-filter frames emitted; *n_out = synthesis frames emitted.
long btkref_chain_zelinski(const btkref_chain_cfg* cfg, const float* pcm, long T, const double* h, const double* g,
                           const double* delays, double alpha, int type, int min_frames, double* Ypf, double* Wpf,
                           long cap, float* out, long cap_out, long* n_out) {
  try {
    const int M = cfg->M, m = cfg->m, r = cfg->r, C = cfg->C;
    const unsigned D = M >> r;
    gsl_vector* hp = make_vector(h, (size_t)M * m);
    SubbandDSPtr bf(new SubbandDS(M, false));
    for (int c = 0; c < C; c++) {
      VectorFloatFeatureStreamPtr src(new MemorySampleFeature(pcm + c, T, C, D));
      VectorComplexFeatureStreamPtr bank(new OverSampledDFTAnalysisBank(src, hp, M, m, r, cfg->dct));
      bf->setChannel(bank);
    }
    gsl_vector_free(hp);
    gsl_vector* dv = make_vector(delays, C);
    bf->calcArrayManifoldVectors(cfg->fs, dv);
    gsl_vector_free(dv);
    ZelinskiPostFilterPtr pf(new ZelinskiPostFilter((VectorComplexFeatureStreamPtr&)bf, M, alpha, type, min_frames));
    pf->setBeamformer(bf);
    long nsyn = 0, nf = 0;
    // a tee is not needed: the post-filter's own buffer is read after every pull
    struct Rec : public VectorComplexFeatureStream {
      ZelinskiPostFilterPtr pf; double* Ypf; double* Wpf; long cap; long n;
      Rec(ZelinskiPostFilterPtr& p, unsigned M, double* y, double* w, long c)
          : VectorComplexFeatureStream(M, "Rec"), pf(p), Ypf(y), Wpf(w), cap(c), n(0) {}
      virtual const gsl_vector_complex* next(int frameX = -5) {
        if (frameX == _frameX) return _vector;
        const gsl_vector_complex* y = pf->next(frameX);
        const unsigned M = size();
        gsl_vector_complex_memcpy(_vector, y);
        if (n < cap) {
          if (Ypf) for (unsigned s = 0; s < M; s++) {
            gsl_complex z = gsl_vector_complex_get(y, s);
            Ypf[2 * (n * M + s)] = GSL_REAL(z); Ypf[2 * (n * M + s) + 1] = GSL_IMAG(z);
          }
          const gsl_vector_complex* w = pf->getPostFilterWeights();
          if (Wpf && w) for (unsigned s = 0; s < M; s++) Wpf[n * M + s] = GSL_REAL(gsl_vector_complex_get(w, s));
        }
        n++;
        _increment();
        return _vector;
      }
      virtual void reset() { pf->reset(); VectorComplexFeatureStream::reset(); }
    };
    Rec* rec = new Rec(pf, M, Ypf, Wpf, cap);
    VectorComplexFeatureStreamPtr recp(rec);
    if (g) {
      gsl_vector* gp = make_vector(g, (size_t)M * m);
      OverSampledDFTSynthesisBankPtr syn(new OverSampledDFTSynthesisBank(recp, gp, M, m, r, cfg->dct, cfg->gain));
      gsl_vector_free(gp);
      try {
        for (;;) {
          const gsl_vector_float* f = syn->next();
          if (out && nsyn < cap_out) for (unsigned d = 0; d < D; d++) out[nsyn * D + d] = gsl_vector_float_get(f, d);
          nsyn++;
        }
      } catch (jiterator_error&) {}
    } else {
      try { for (;;) recp->next(); } catch (jiterator_error&) {}
    }
    if (n_out) *n_out = nsyn;
    nf = rec->n;
    return nf;
  } catch (std::exception& e) { fprintf(stderr, "btkref_chain_zelinski: %s\n", e.what()); return -1; }
}

// analysis banks -> SubbandGSC (calcGSCWeights, then setActiveWeights_f for every bin 0..M/2) -> (optional) synthesis.
//   wa : [B][C-1][2] active weights;  Bout: [B][C][C-1][2] blocking matrices or NULL;  wq: [B][C][2] or NULL
long btkref_chain_gsc(const btkref_chain_cfg* cfg, const float* pcm, long T, const double* h, const double* g,
                      const double* delays, const double* wa, int normalize, double* Y, long cap, float* out, long cap_out,
                      long* n_out, double* Bout, double* wq) {
  try {
    const int M = cfg->M, m = cfg->m, r = cfg->r, C = cfg->C;
    const unsigned D = M >> r, B = M / 2 + 1;
    gsl_vector* hp = make_vector(h, (size_t)M * m);
    SubbandGSC* bf = new SubbandGSC(M, false);
    for (int c = 0; c < C; c++) {
      VectorFloatFeatureStreamPtr src(new MemorySampleFeature(pcm + c, T, C, D));
      VectorComplexFeatureStreamPtr bank(new OverSampledDFTAnalysisBank(src, hp, M, m, r, cfg->dct));
      bf->setChannel(bank);
    }
    gsl_vector_free(hp);
    gsl_vector* dv = make_vector(delays, C);
    bf->calcGSCWeights(cfg->fs, dv);
    gsl_vector_free(dv);
    bf->normalizeWeight(normalize != 0);
    gsl_vector* pw = gsl_vector_alloc(2 * (C - 1));
    for (unsigned s = 0; s < B; s++) {
      for (int k = 0; k < 2 * (C - 1); k++) gsl_vector_set(pw, k, wa[(size_t)s * 2 * (C - 1) + k]);
      bf->setActiveWeights_f(s, pw);
      if (Bout) {
        const gsl_matrix_complex* Bm = bf->getBlockingMatrix(0, s);
        for (int i = 0; i < C; i++) for (int k = 0; k < C - 1; k++) {
          gsl_complex z = gsl_matrix_complex_get(Bm, i, k);
          double* p = Bout + 2 * (((size_t)s * C + i) * (C - 1) + k);
          p[0] = GSL_REAL(z); p[1] = GSL_IMAG(z);
        }
      }
      if (wq) {
        const gsl_vector_complex* w = bf->getWeights(s);
        for (int c = 0; c < C; c++) {
          gsl_complex z = gsl_vector_complex_get(w, c);
          wq[2 * (s * C + c)] = GSL_REAL(z); wq[2 * (s * C + c) + 1] = GSL_IMAG(z);
        }
      }
    }
    gsl_vector_free(pw);
    RecordingTee* tee = new RecordingTee(bf, M, C, Y, NULL, cap);
    VectorComplexFeatureStreamPtr teep(tee);
    long nsyn = 0;
    if (g) {
      gsl_vector* gp = make_vector(g, (size_t)M * m);
      OverSampledDFTSynthesisBankPtr syn(new OverSampledDFTSynthesisBank(teep, gp, M, m, r, cfg->dct, cfg->gain));
      gsl_vector_free(gp);
      try {
        for (;;) {
          const gsl_vector_float* f = syn->next();
          if (out && nsyn < cap_out) for (unsigned d = 0; d < D; d++) out[nsyn * D + d] = gsl_vector_float_get(f, d);
          nsyn++;
        }
      } catch (jiterator_error&) {}
    } else {
      try { for (;;) teep->next(); } catch (jiterator_error&) {}
    }
    if (n_out) *n_out = nsyn;
    long nf = tee->frames();
    bf->clearChannel();
    delete bf;
    return nf;
  } catch (std::exception& e) { fprintf(stderr, "btkref_chain_gsc: %s\n", e.what()); return -1; }
}

// de Haan prototype design (prototypeDesign.cc:611-951): h = pinv(A + C) b, then g = pinv(E + v P) f from that h.
//   h, g: [M*m] outputs;  err_h[3] / err_g[3]: calcError() of the two designs (dB).  Returns 0, <0 on error.
int btkref_design_dehaan(int M, int m, int r, double wpFactor, double v, double tolerance, double* h, double* g,
                         double* err_h, double* err_g) {
  try {
    const int L = M * m;
    AnalysisOversampledDFTDesign ana(M, m, r, wpFactor, -1);
    const gsl_vector* hv = ana.design(tolerance);
    for (int n = 0; n < L; n++) h[n] = gsl_vector_get(hv, n);
    const gsl_vector* eh = ana.calcError(false);
    if (err_h) for (int k = 0; k < 3; k++) err_h[k] = gsl_vector_get(eh, k);
    if (g) {
      gsl_vector* hc = make_vector(h, L);
      SynthesisDesignWithWorkspace syn(hc, M, m, r, v, wpFactor, -1);
      const gsl_vector* gv = syn.design(tolerance);
      for (int n = 0; n < L; n++) g[n] = gsl_vector_get(gv, n);
      const gsl_vector* eg = syn.calcError(false);
      if (err_g) for (int k = 0; k < 3; k++) err_g[k] = gsl_vector_get(eg, k);
      gsl_vector_free(hc);
    }
    return 0;
  } catch (std::exception& e) { fprintf(stderr, "btkref_design_dehaan: %s\n", e.what()); return -1; }
}

// SpectralMatrixArray recursion (beamformer.cc:142-163) over all frames of a recording.
// Rout: [M][C][C][2].  Returns frames consumed.
long btkref_spectral_matrix(const float* pcm, long T, int C, const double* h, int M, int m, int r, int dct,
                            double mu, double* Rout) {
  try {
    const unsigned D = M >> r;
    gsl_vector* hp = make_vector(h, (size_t)M * m);
    std::vector<VectorComplexFeatureStreamPtr> banks;
    for (int c = 0; c < C; c++) {
      VectorFloatFeatureStreamPtr src(new MemorySampleFeature(pcm + c, T, C, D));
      banks.push_back(VectorComplexFeatureStreamPtr(new OverSampledDFTAnalysisBank(src, hp, M, m, r, dct)));
    }
    gsl_vector_free(hp);
    SpectralMatrixArray sma(M, C, mu);
    sma.zero();
    long n = 0;
    try {
      for (;;) {
        for (int c = 0; c < C; c++) sma.newSample(banks[c]->next(), c);
        sma.update();
        n++;
      }
    } catch (jiterator_error&) {}
    for (int s = 0; s < M; s++) {
      const gsl_matrix_complex* Rm = sma.getSpecMatrix(s);
      for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) {
        gsl_complex z = gsl_matrix_complex_get(Rm, i, j);
        double* p = Rout + 2 * ((size_t)(s * C + i) * C + j);
        p[0] = GSL_REAL(z); p[1] = GSL_IMAG(z);
      }
    }
    return n;
  } catch (std::exception& e) { fprintf(stderr, "btkref_spectral_matrix: %s\n", e.what()); return -1; }
}

// Error-path probes used by the parity tests of the drop-in classes' exception behaviour.
//   which: 0 = prototype size mismatch (modulated.cc:269-271)  -> expect jconsistency_error
//          1 = delays size mismatch   (beamformer.cc:533-535)  -> expect jdimension_error
//          2 = SubbandDS::next before weights (beamformer.cc:1140-1143) -> j_error
// Returns the j_error code thrown, or -1 when nothing was thrown.
int btkref_error_probe(int which) {
  try {
    if (which == 0) {
      gsl_vector* p = gsl_vector_calloc(10);
      float x[8] = {0};
      VectorFloatFeatureStreamPtr src(new MemorySampleFeature(x, 8, 1, 4));
      OverSampledDFTAnalysisBankPtr bank(new OverSampledDFTAnalysisBank(src, p, 8, 2, 1, 0));
    } else if (which == 1) {
      SubbandDS bf(8, false);
      gsl_vector* p = gsl_vector_calloc(16);
      float x[8] = {0};
      VectorFloatFeatureStreamPtr src(new MemorySampleFeature(x, 8, 1, 4));
      VectorComplexFeatureStreamPtr bank(new OverSampledDFTAnalysisBank(src, p, 8, 2, 1, 0));
      bf.setChannel(bank);
      gsl_vector* d = gsl_vector_calloc(3);
      bf.calcArrayManifoldVectors(16000.0, d);
    } else if (which == 2) {
      SubbandDS bf(8, false);
      bf.next();
    } else if (which == 3) {
      // end of stream: the reference's own analysis bank throws jiterator_error once its pd padded frames are out
      // (modulated.cc:503-515)
      gsl_vector* p = gsl_vector_calloc(16);
      float x[8] = {0};
      VectorFloatFeatureStreamPtr src(new MemorySampleFeature(x, 8, 1, 4));
      OverSampledDFTAnalysisBankPtr bank(new OverSampledDFTAnalysisBank(src, p, 8, 2, 1, 0));
      for (int i = 0; i < 64; i++) bank->next();
    }
  } catch (j_error& e) { return (int)e.getCode(); }
  catch (std::exception&) { return -2; }
  return -1;
}

// Quiescent weights of SubbandDS::calcArrayManifoldVectorsN (beamformer.cc:1113-1121 -> calcMainlobeN :632-735) for bins
// 0..M/2, and the array manifold the weight object keeps next to them.  delaysJ: [NC-1][C].  w, ta: [M/2+1][C][2].
int btkref_null_weights(double fs, const double* delaysT, const double* delaysJ, int M, int C, int NC, double* w, double* ta) {
  try {
    SubbandDS bf(M, false);
    static double none = 0;
    for (int c = 0; c < C; c++) {
      VectorComplexFeatureStreamPtr ch(new MemoryComplexFeature(&none, 0, M));
      bf.setChannel(ch);
    }
    gsl_vector* dT = make_vector(delaysT, C);
    gsl_matrix* dJ = gsl_matrix_alloc(NC - 1, C);
    for (int n = 0; n < NC - 1; n++) for (int c = 0; c < C; c++) gsl_matrix_set(dJ, n, c, delaysJ[(size_t)n * C + c]);
    if (NC == 2) {
      gsl_vector* dj = make_vector(delaysJ, C);
      bf.calcArrayManifoldVectors2(fs, dT, dj);
      gsl_vector_free(dj);
    } else {
      bf.calcArrayManifoldVectorsN(fs, dT, dJ, NC);
    }
    beamformerWeights* bw = bf.getBeamformerWeightObject(0);
    for (int s = 0; s <= M / 2; s++)
      for (int c = 0; c < C; c++) {
        gsl_complex z = gsl_vector_complex_get(bw->wq_f(s), c), t = gsl_vector_complex_get(bw->arrayManifold()[s], c);
        w[2 * ((size_t)s * C + c)] = GSL_REAL(z); w[2 * ((size_t)s * C + c) + 1] = GSL_IMAG(z);
        ta[2 * ((size_t)s * C + c)] = GSL_REAL(t); ta[2 * ((size_t)s * C + c) + 1] = GSL_IMAG(t);
      }
    gsl_vector_free(dT); gsl_matrix_free(dJ);
    return 0;
  } catch (std::exception& e) { fprintf(stderr, "btkref_null_weights: %s\n", e.what()); return -1; }
}

// calcAllDelays (beamformer.cc:1214-1231) as compiled from the reference.
int btkref_calc_all_delays(double x, double y, double z, const double* micpos, int n, double* delays) {
  gsl_matrix* mp = gsl_matrix_alloc(n, 3);
  for (int c = 0; c < n; c++) for (int k = 0; k < 3; k++) gsl_matrix_set(mp, c, k, micpos[3 * c + k]);
  gsl_vector* d = gsl_vector_alloc(n);
  calcAllDelays(x, y, z, mp, d);
  for (int c = 0; c < n; c++) delays[c] = gsl_vector_get(d, c);
  gsl_vector_free(d); gsl_matrix_free(mp);
  return 0;
}

// calcDelaysPolar2 lives in a driver translation unit with a main() (src/superdirectiveBeamformer.cc:118-137); oracle/Makefile
// extracts exactly those lines (and the SOUNDSPEED definition, :14) from the reference source into
// _ref/obj/calc_delays_polar2.inc at build time, so the function below is the reference's text, compiled -- not a copy kept
// in this repository.
}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// Ingest nodes: IterativeSampleFeature::next (feature/feature.cc:868-896), Conversion24bit2Float::next
// (:190-217), ChannelExtractionFeature::next (:3885-3900).  feature.cc itself needs libsndfile and much of GSL, so
// oracle/Makefile extracts exactly those three definitions into _ref/obj/ingest_next.inc; they are compiled here, verbatim,
// as members of classes that DECLARE the same members as feature/feature.h:148-158, 301-335, 1823-1841 (constructors and
// read() restated from :803-866: the interval buffer of 30 s, _blockN = interval * samplerate / blockLen + 1).  The only
// stand-in is sf_readf_float, which reads interleaved floats from memory (libsndfile with SFC_SET_NORM_FLOAT off hands
// the integer sample values over unscaled, :849).
namespace ingest_ref {
namespace sndfile {
struct SNDFILE { const float* data; long frames; long pos; int channels; };
struct SF_INFO { long frames; int samplerate, channels, format; };
}
static unsigned sf_readf_float(sndfile::SNDFILE* f, float* dst, unsigned n) {
  long left = f->frames - f->pos;
  unsigned got = left <= 0 ? 0 : (left < (long)n ? (unsigned)left : n);
  memcpy(dst, f->data + f->pos * f->channels, sizeof(float) * (size_t)got * f->channels);
  f->pos += got;
  return got;
}

class IterativeSampleFeature : public VectorFloatFeatureStream {
 public:
  IterativeSampleFeature(unsigned chX, unsigned blockLen = 320, unsigned firstChanX = 0, const String& nm = "Iterative Sample")
      : VectorFloatFeatureStream(blockLen, nm), _blockLen(blockLen), _chanX(chX), _firstChanX(firstChanX), _cur(0) {}
  // read() of the reference with the file replaced by a memory image (:826-866)
  void read_memory(const float* data, long frames, int samplerate, int chN, int cfrom, int cto) {
    if (_chanX != _firstChanX) return;
    delete[] _allSamples; _allSamples = NULL;
    delete _sndfile;
    _sfinfo.channels = chN; _sfinfo.samplerate = samplerate; _sfinfo.format = 0;
    _sndfile = new sndfile::SNDFILE();
    _sndfile->data = data; _sndfile->frames = frames; _sndfile->pos = 0; _sndfile->channels = chN;
    _blockN = _interval * _sfinfo.samplerate / _blockLen + 1;
    _sampleN = _blockN * _blockLen;
    _allSampleN = _sampleN * _sfinfo.channels;
    _allSamples = new float[_allSampleN];
    _sndfile->pos = cfrom;
    _cto = cto - cfrom;
  }
  unsigned samplesN() const { return _ttlSamples; }
  virtual const gsl_vector_float* next(int frameX = -5);
  virtual void reset() { _ttlSamples = _cur = 0; _last = false; VectorFloatFeatureStream::reset(); }

 private:
  static float* _allSamples;
  static sndfile::SNDFILE* _sndfile;
  static sndfile::SF_INFO _sfinfo;
  static unsigned _interval, _blockN, _sampleN, _allSampleN, _ttlSamples;
  const unsigned _blockLen;
  const unsigned _chanX;
  unsigned _firstChanX;
  unsigned _cur;
  bool _last;
  int _cto;
};
float* IterativeSampleFeature::_allSamples = NULL;
sndfile::SNDFILE* IterativeSampleFeature::_sndfile = NULL;
sndfile::SF_INFO IterativeSampleFeature::_sfinfo;
unsigned IterativeSampleFeature::_interval = 30;
unsigned IterativeSampleFeature::_blockN;
unsigned IterativeSampleFeature::_sampleN;
unsigned IterativeSampleFeature::_allSampleN;
unsigned IterativeSampleFeature::_ttlSamples;

class Conversion24bit2Float : public VectorFloatFeatureStream {
 public:
  Conversion24bit2Float(VectorCharFeatureStreamPtr& src, const String& nm = "Conversion from 24 bit integer to Float")
      : VectorFloatFeatureStream(src->size() / 3, nm), _src(src) {}
  virtual void reset() { _src->reset(); VectorFloatFeatureStream::reset(); }
  virtual const gsl_vector_float* next(int frameX = -5);
 private:
  VectorCharFeatureStreamPtr _src;
};

class ChannelExtractionFeature : public VectorFloatFeatureStream {
 public:
  ChannelExtractionFeature(const VectorFloatFeatureStreamPtr& src, unsigned chX = 0, unsigned chN = 1, const String& nm = "ChannelExtraction")
      : VectorFloatFeatureStream(src->size() / chN, nm), _src(src), _chX(chX), _chN(chN) {}
  virtual const gsl_vector_float* next(int frameX = -5);
  virtual void reset() { _src->reset(); VectorFloatFeatureStream::reset(); }
 private:
  VectorFloatFeatureStreamPtr _src;
  unsigned _chX, _chN;
};

#include "ingest_next.inc"

// block sources over memory for the two wrappers
class MemoryCharFeature : public VectorCharFeatureStream {
 public:
  MemoryCharFeature(const char* data, long n, unsigned block) : VectorCharFeatureStream(block, "MemoryChar"), _d(data), _n(n), _cur(0) {}
  virtual const gsl_vector_char* next(int frameX = -5) {
    if (frameX == _frameX) return _vector;
    if (_cur + (long)size() > _n) throw jiterator_error("end of samples!");
    for (unsigned i = 0; i < size(); i++) gsl_vector_char_set(_vector, i, _d[_cur + i]);
    _cur += size();
    _increment();
    return _vector;
  }
  virtual void reset() { _cur = 0; VectorCharFeatureStream::reset(); }
 private:
  const char* _d; long _n, _cur;
};
class MemoryFloatBlocks : public VectorFloatFeatureStream {
 public:
  MemoryFloatBlocks(const float* data, long n, unsigned block) : VectorFloatFeatureStream(block, "MemoryFloat"), _d(data), _n(n), _cur(0) {}
  virtual const gsl_vector_float* next(int frameX = -5) {
    if (frameX == _frameX) return _vector;
    if (_cur + (long)size() > _n) throw jiterator_error("end of samples!");
    for (unsigned i = 0; i < size(); i++) gsl_vector_float_set(_vector, i, _d[_cur + i]);
    _cur += size();
    _increment();
    return _vector;
  }
  virtual void reset() { _cur = 0; VectorFloatFeatureStream::reset(); }
 private:
  const float* _d; long _n, _cur;
};
}  // namespace ingest_ref

extern "C" {
// All channels of an interleaved recording through chN IterativeSampleFeature nodes pulled in lock step, the way the
// multichannel drivers do (channel `firstChanX` first).  out: [cap_blocks][chN][blockLen]; returns blocks served per channel.
long btkref_iterative_sample(const float* data, long frames, int samplerate, int chN, int blockLen, int cfrom, int cto,
                             float* out, long cap_blocks, long* ttl_samples) {
  try {
    std::vector<VectorFloatFeatureStreamPtr> nodes;
    std::vector<ingest_ref::IterativeSampleFeature*> raw;
    for (int c = 0; c < chN; c++) {
      ingest_ref::IterativeSampleFeature* f = new ingest_ref::IterativeSampleFeature(c, blockLen, 0);
      raw.push_back(f); nodes.push_back(VectorFloatFeatureStreamPtr(f));
    }
    for (int c = 0; c < chN; c++) { raw[c]->reset(); raw[c]->read_memory(data, frames, samplerate, chN, cfrom, cto); }
    long n = 0;
    try {
      for (;;) {
        for (int c = 0; c < chN; c++) {
          const gsl_vector_float* b = nodes[c]->next();
          if (n < cap_blocks) for (int i = 0; i < blockLen; i++) out[((size_t)n * chN + c) * blockLen + i] = gsl_vector_float_get(b, i);
        }
        n++;
        if (n > cap_blocks + 8) break;
      }
    } catch (jiterator_error&) {}
    if (ttl_samples) *ttl_samples = raw[0]->samplesN();
    return n;
  } catch (std::exception& e) { fprintf(stderr, "btkref_iterative_sample: %s\n", e.what()); return -1; }
}

// Conversion24bit2Float over a byte stream cut into blocks of 3 * block bytes.  Returns samples written.
long btkref_conversion24(const char* bytes, long nbytes, int block, float* out) {
  try {
    VectorCharFeatureStreamPtr src(new ingest_ref::MemoryCharFeature(bytes, nbytes, 3 * block));
    ingest_ref::Conversion24bit2Float conv(src);
    long n = 0;
    try { for (;;) { const gsl_vector_float* b = conv.next(); for (int i = 0; i < block; i++) out[n++] = gsl_vector_float_get(b, i); } }
    catch (jiterator_error&) {}
    return n;
  } catch (std::exception& e) { fprintf(stderr, "btkref_conversion24: %s\n", e.what()); return -1; }
}

// ChannelExtractionFeature(chX of chN) over interleaved blocks of chN * block floats.  Returns samples written.
long btkref_channel_extraction(const float* data, long n, int chX, int chN, int block, float* out) {
  try {
    VectorFloatFeatureStreamPtr src(new ingest_ref::MemoryFloatBlocks(data, n, chN * block));
    ingest_ref::ChannelExtractionFeature ext(src, chX, chN);
    long k = 0;
    try { for (;;) { const gsl_vector_float* b = ext.next(); for (int i = 0; i < block; i++) out[k++] = gsl_vector_float_get(b, i); } }
    catch (jiterator_error&) {}
    return k;
  } catch (std::exception& e) { fprintf(stderr, "btkref_channel_extraction: %s\n", e.what()); return -1; }
}
}  // extern "C"
#include "calc_delays_polar2.inc"
extern "C" {
int btkref_calc_delays_polar2(float azimuth, float elevation, const double* micpos, int n, double* delays) {
  gsl_matrix* mp = gsl_matrix_alloc(n, 3);
  for (int c = 0; c < n; c++) for (int k = 0; k < 3; k++) gsl_matrix_set(mp, c, k, micpos[3 * c + k]);
  gsl_vector* d = calcDelaysPolar2(azimuth, elevation, mp);
  for (int c = 0; c < n; c++) delays[c] = gsl_vector_get(d, c);
  gsl_vector_free(d); gsl_matrix_free(mp);
  return 0;
}

}  // extern "C"
