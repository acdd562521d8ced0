// btk_streams.h -- drop-in C++ feature-stream nodes for the hot path, above the C ABI (include/btkb200.h).
//
// Same class names, constructor arguments, next()/reset() semantics and exception types as the reference
// (paths relative to /root/reference/btk):
//   FeatureStream<Type,item>            stream/stream.h:35-75
//   SampleFeature block/pad rule        feature/feature.cc:610-659        (MemorySampleFeature: in-memory source)
//   OverSampledDFTAnalysisBank          modulated/modulated.h:291-322, modulated.cc:359-516
//   OverSampledDFTSynthesisBank         modulated/modulated.h:327-366, modulated.cc:521-674
//   SubbandBeamformer / SubbandDS       beamformer/beamformer.h:126-182, beamformer.cc:1017-1212
//   SubbandMVDR                         beamformer/beamformer.h:333-388, beamformer.cc:2321-2635
//   j_error family                      common/jexception.h:41-173
//
// Header-only host code: it holds no arithmetic of the hot loops.  Nodes evaluate lazily per utterance: the
// first next() drains the upstream node (any VectorFeatureStream, ours or foreign), runs the sm_100a kernels
// through the C ABI once and then serves frames; a synthesis bank sitting on a SubbandDS/SubbandMVDR whose
// channels are all OverSampledDFTAnalysisBank nodes runs the whole chain as ONE fused kernel (btkb200_chain).
// There is no CPU fallback: without a CUDA device the first next() throws j_error with the library's message.
//
// Types.  The reference's element containers are gsl_vector_float / gsl_vector_complex / gsl_vector /
// gsl_matrix(_complex).  GSL is not part of this build, so this header defines containers with the SAME field
// layout (size, stride, data, block, owner) under the names btk_vector_*; inside the reference tree define
// BTKB200_WITH_GSL before including it and the gsl types are used directly.
#ifndef BTKB200_STREAMS_H
#define BTKB200_STREAMS_H

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <exception>
#include <list>
#include <memory>
#include <string>
#include <vector>

#include "../../include/btkb200.h"

#ifdef BTKB200_WITH_GSL
#include <gsl/gsl_matrix.h>
#include <gsl/gsl_vector.h>
typedef gsl_vector_float btk_vector_float;
typedef gsl_vector_complex btk_vector_complex;
typedef gsl_vector btk_vector;
typedef gsl_matrix btk_matrix;
typedef gsl_matrix_complex btk_matrix_complex;
#else
struct btk_vector_float { size_t size, stride; float* data; void* block; int owner; };
struct btk_vector_complex { size_t size, stride; double* data; void* block; int owner; };   // (re, im) pairs
struct btk_vector { size_t size, stride; double* data; void* block; int owner; };
struct btk_matrix { size_t size1, size2, tda; double* data; void* block; int owner; };
struct btk_matrix_complex { size_t size1, size2, tda; double* data; void* block; int owner; };
#endif

namespace btkb200 {

typedef std::string String;

// ---- exceptions (common/jexception.h:41-173: code + formatted message) ---------------------------------------
// error_type of common/jexception.h:41-57, every value in the reference's order (0-based): the SWIG wrappers switch on
// getCode() (include/jexception.i:63-69: JITERATOR -> StopIteration, JIO -> IOError), so the numbers are part of the ABI
enum error_type { JERROR = 0, JALLOCATION, JARITHMETIC, JCONSISTENCY, JDIMENSION, JINDEX, JINITIALIZATION, JIO, JITERATOR,
                  JPYTHON, JKEY, JNUMERIC, JPARAMETER, JPARSE, JTYPE };
static_assert(JCONSISTENCY == 3 && JDIMENSION == 4 && JIO == 7 && JITERATOR == 8 && JPYTHON == 9, "common/jexception.h:41-57");
class j_error : public std::exception {
 public:
  j_error() : _code(JERROR) {}
  explicit j_error(const char* fmt, ...) : _code(JERROR) { va_list ap; va_start(ap, fmt); format(fmt, ap); va_end(ap); }
  virtual ~j_error() throw() {}
  virtual const char* what() const throw() { return _what.c_str(); }
  int getCode() const { return _code; }
 protected:
  void format(const char* fmt, va_list ap) { char b[512]; vsnprintf(b, sizeof b, fmt, ap); _what = b; }
  int _code;
  std::string _what;
};
#define BTKB200_DEFINE_ERROR(NAME, CODE)                                                              \
  class NAME : public j_error {                                                                       \
   public:                                                                                            \
    explicit NAME(const char* fmt, ...) { _code = CODE; va_list ap; va_start(ap, fmt); format(fmt, ap); va_end(ap); } \
  };
BTKB200_DEFINE_ERROR(jiterator_error, JITERATOR)       // doubles as end of stream (-> StopIteration in Python)
BTKB200_DEFINE_ERROR(jconsistency_error, JCONSISTENCY)
BTKB200_DEFINE_ERROR(jdimension_error, JDIMENSION)
#undef BTKB200_DEFINE_ERROR

// status of the C ABI -> the exception the reference throws at that point (include/btkb200.h)
inline void check(int rc, const btkb200_plan* plan) {
  if (rc == BTKB200_OK) return;
  const char* msg = btkb200_last_error(plan);
  if (rc == BTKB200_EINVAL) throw jdimension_error("%s", msg);
  throw j_error("%s", msg);
}

// ---- FeatureStream (stream/stream.h:35-75) ------------------------------------------------------------------------
template <typename Type, typename item_type>
class FeatureStream {
 public:
  virtual ~FeatureStream() { delete[] _vector->data; delete _vector; }
  const String& name() const { return _name; }
  unsigned size() const { return _size; }
  virtual const Type* next(int frameX = -5) = 0;
  const Type* current() {
    if (_frameX < 0) throw jconsistency_error("Frame index (%d) < 0.", _frameX);
    return next(_frameX);
  }
  bool isEnd() { return _endOfSamples; }
  virtual void reset() { _frameX = FrameResetX; _endOfSamples = false; }
  virtual int frameX() const { return _frameX; }

 protected:
  FeatureStream(unsigned sz, const String& nm, unsigned items_per_element)
      : FrameResetX(-1), _size(sz), _frameX(-1), _endOfSamples(false), _name(nm) {
    _vector = new Type();
    _vector->size = sz; _vector->stride = 1; _vector->block = 0; _vector->owner = 0;
    _vector->data = new item_type[(size_t)sz * items_per_element]();
  }
  void _increment() { _frameX++; }
  const int FrameResetX;
  const unsigned _size;
  int _frameX;
  Type* _vector;
  bool _endOfSamples;

 private:
  const String _name;
};

class VectorFloatFeatureStream : public FeatureStream<btk_vector_float, float> {
 protected:
  VectorFloatFeatureStream(unsigned sz, const String& nm) : FeatureStream<btk_vector_float, float>(sz, nm, 1) {}
};
class VectorComplexFeatureStream : public FeatureStream<btk_vector_complex, double> {
 protected:
  VectorComplexFeatureStream(unsigned sz, const String& nm) : FeatureStream<btk_vector_complex, double>(sz, nm, 2) {}
};
typedef std::shared_ptr<VectorFloatFeatureStream> VectorFloatFeatureStreamPtr;
typedef std::shared_ptr<VectorComplexFeatureStream> VectorComplexFeatureStreamPtr;

// ---- block source with SampleFeature::next's rule (feature/feature.cc:610-659) ----------------------------------
class MemorySampleFeature : public VectorFloatFeatureStream {
 public:
  MemorySampleFeature(const float* samples, size_t n, unsigned blockLen, unsigned shiftLen, bool padZeros,
                      const String& nm = "Sample")
      : VectorFloatFeatureStream(blockLen, nm), _samples(samples, samples + n), _shift(shiftLen), _pad(padZeros), _cur(0) {}
  virtual const btk_vector_float* next(int frameX = -5) {
    if (frameX == _frameX) return _vector;
    if (frameX >= 0 && frameX - 1 != _frameX)
      throw jconsistency_error("Problem in Feature %s: %d != %d", name().c_str(), frameX - 1, _frameX);
    const size_t T = _samples.size();
    if (_cur >= T || (!_pad && _cur + _size > T)) { _endOfSamples = true; throw jiterator_error("end of samples!"); }
    for (unsigned i = 0; i < _size; i++) _vector->data[i] = _cur + i < T ? _samples[_cur + i] : 0.f;
    _cur += _shift;
    _increment();
    return _vector;
  }
  virtual void reset() { VectorFloatFeatureStream::reset(); _cur = 0; }
  const std::vector<float>& samples() const { return _samples; }
  unsigned shiftLen() const { return _shift; }
  bool padZeros() const { return _pad; }

 private:
  std::vector<float> _samples;
  unsigned _shift;
  bool _pad;
  size_t _cur;
};

// owning handle of one btkb200_plan
class PlanHandle {
 public:
  PlanHandle() : _p(0) {}
  ~PlanHandle() { reset(); }
  void reset() { if (_p) btkb200_plan_destroy(_p); _p = 0; }
  void create(unsigned M, unsigned m, unsigned r, unsigned dct, unsigned C, const double* h, const double* g, int gain) {
    reset();
    const int rc = btkb200_plan_create(&_p, M, m, r, dct, C, h, g, gain, 0);
    if (rc != BTKB200_OK) { _p = 0; check(rc, 0); }
  }
  btkb200_plan* get() const { return _p; }
  operator bool() const { return _p != 0; }
  unsigned C() const { btkb200_info i; btkb200_plan_info(_p, &i); return i.C; }
  int has_weights() const { btkb200_info i; btkb200_plan_info(_p, &i); return i.has_weights; }

 private:
  PlanHandle(const PlanHandle&);
  PlanHandle& operator=(const PlanHandle&);
  btkb200_plan* _p;
};

// ---- OverSampledDFTFilterBank geometry (modulated/modulated.cc:262-300) -------------------------------------------
class OverSampledDFTFilterBank {
 public:
  double polyphase(unsigned m, unsigned n) const { return _prototype[m + _M * n]; }   // modulated.h:233-236
 protected:
  OverSampledDFTFilterBank(const btk_vector* prototype, unsigned M, unsigned m, unsigned r, unsigned dct)
      : _M(M), _m(m), _r(r), _R(1u << r), _D(M >> r), _dct(dct) {
    if (!prototype || prototype->size != (size_t)M * m)
      throw jconsistency_error("Prototype sizes do not match (%d vs. %d).", prototype ? (int)prototype->size : 0, (int)(M * m));
    _prototype.resize((size_t)M * m);
    for (size_t i = 0; i < _prototype.size(); i++) _prototype[i] = prototype->data[i * prototype->stride];   // copied (:275-276)
  }
 public:
  const unsigned _M, _m, _r, _R, _D, _dct;
  std::vector<double> _prototype;
};

// ---- OverSampledDFTAnalysisBank ---------------------------------------------------------------------------------
class OverSampledDFTAnalysisBank : public OverSampledDFTFilterBank, public VectorComplexFeatureStream {
 public:
  OverSampledDFTAnalysisBank(VectorFloatFeatureStreamPtr& samp, btk_vector* prototype, unsigned M, unsigned m, unsigned r,
                             unsigned delayCompensationType = 0, const String& nm = "OverSampledDFTAnalysisBank")
      : OverSampledDFTFilterBank(prototype, M, m, r, delayCompensationType), VectorComplexFeatureStream(M, nm), _samp(samp),
        _F(-1) {
    if (samp->size() != _D) throw jdimension_error("Input block length (%d) != D (%d)\n", samp->size(), _D);   // :373-374
  }
  unsigned fftLen() const { return _M; }
  unsigned nBlocks() const { return 4; }
  unsigned subSampRate() const { return 2; }

  virtual const btk_vector_complex* next(int frameX = -5) {
    if (frameX == _frameX && _frameX >= 0) return _vector;
    if (_F < 0) evaluate();
    const int t = _frameX + 1;
    if (t >= _F) { _endOfSamples = true; throw jiterator_error("end of samples!"); }
    const unsigned B = _M / 2 + 1;
    const float* half = &_frames[(size_t)t * B * 2];
    for (unsigned s = 0; s < B; s++) { _vector->data[2 * s] = half[2 * s]; _vector->data[2 * s + 1] = half[2 * s + 1]; }
    for (unsigned s = 1; s < _M / 2; s++) { _vector->data[2 * (_M - s)] = half[2 * s]; _vector->data[2 * (_M - s) + 1] = -half[2 * s + 1]; }
    _increment();
    return _vector;
  }
  virtual void reset() { _samp->reset(); VectorComplexFeatureStream::reset(); _F = -1; }

  // every sample of the source (the analysis bank swallows any exception from it and pads, modulated.cc:493-501)
  void pull_source(std::vector<float>& x) {
    x.clear();
    MemorySampleFeature* ms = dynamic_cast<MemorySampleFeature*>(_samp.get());
    if (ms && ms->shiftLen() == _D && ms->padZeros() && ms->frameX() < 0) { x = ms->samples(); return; }
    for (;;) {
      const btk_vector_float* b;
      try { b = _samp->next(); } catch (std::exception&) { break; }
      for (unsigned i = 0; i < _D; i++) x.push_back(b->data[i * b->stride]);
    }
  }

 private:
  void evaluate() {
    std::vector<float> x;
    pull_source(x);
    if (!_plan) _plan.create(_M, _m, _r, _dct, 1, &_prototype[0], 0, 1);
    const long F = btkb200_analysis_frames(_plan.get(), (long)x.size());
    _frames.assign((size_t)F * (_M / 2 + 1) * 2, 0.f);
    long n = 0;
    float dummy = 0.f;
    check(btkb200_analysis(_plan.get(), x.empty() ? &dummy : &x[0], (long)x.size(), &_frames[0], &n), _plan.get());
    _F = (int)n;
  }
  VectorFloatFeatureStreamPtr _samp;
  PlanHandle _plan;
  std::vector<float> _frames;   // [F][B] complex64
  int _F;
};
typedef std::shared_ptr<OverSampledDFTAnalysisBank> OverSampledDFTAnalysisBankPtr;

// ---- SubbandBeamformer / SubbandDS / SubbandMVDR ----------------------------------------------------------------
class SubbandBeamformer : public VectorComplexFeatureStream {
 public:
  SubbandBeamformer(unsigned fftLen = 512, bool halfBandShift = false, const String& nm = "SubbandBeamformer")
      : VectorComplexFeatureStream(fftLen, nm), _fftLen(fftLen), _fftLen2(fftLen / 2), _F(-1) {
    if (halfBandShift) throw j_error("halfBandShift is not supported by the B200 engine");
    _snap = new btk_vector_complex(); _snap->size = 0; _snap->stride = 1; _snap->data = 0; _snap->block = 0; _snap->owner = 0;
    _wvec = new btk_vector_complex(); *_wvec = *_snap;
  }
  ~SubbandBeamformer() { delete[] _snap->data; delete _snap; delete[] _wvec->data; delete _wvec; }
  unsigned fftLen() const { return _fftLen; }
  unsigned fftLen2() const { return _fftLen2; }
  unsigned chanN() const { return (unsigned)_channelList.size(); }
  virtual unsigned dim() const { return chanN(); }
  void setChannel(VectorComplexFeatureStreamPtr& chan) { _channelList.push_back(chan); }
  virtual void clearChannel() { _channelList.clear(); _plan.reset(); _F = -1; }

  // snapshot of the CURRENT frame at one bin, C complex values (SnapShotArray::getSnapShot)
  const btk_vector_complex* snapShotArray_f(unsigned fbinX) {
    if (_F < 0 || _frameX < 0) throw j_error("no snapshot yet");
    const unsigned C = chanN(), B = _fftLen / 2 + 1;
    fill(_snap, C);
    const float* s = &_snapshots[(((size_t)_frameX * B) + fbinX) * C * 2];
    for (unsigned i = 0; i < 2 * C; i++) _snap->data[i] = s[i];
    return _snap;
  }

  virtual const btk_vector_complex* next(int frameX = -5) {
    if (frameX == _frameX && _frameX >= 0) return _vector;
    if (_F < 0) evaluate();
    const int t = _frameX + 1;
    if (t >= _F) { _endOfSamples = true; throw jiterator_error("end of samples!"); }
    const unsigned B = _fftLen / 2 + 1, M = _fftLen;
    const float* half = &_Y[(size_t)t * B * 2];
    for (unsigned s = 0; s < B; s++) { _vector->data[2 * s] = half[2 * s]; _vector->data[2 * s + 1] = half[2 * s + 1]; }
    for (unsigned s = 1; s < M / 2; s++) { _vector->data[2 * (M - s)] = half[2 * s]; _vector->data[2 * (M - s) + 1] = -half[2 * s + 1]; }   // :1189-1194
    _increment();
    return _vector;
  }
  virtual void reset() {
    for (_ChannelIterator it = _channelList.begin(); it != _channelList.end(); ++it) (*it)->reset();
    VectorComplexFeatureStream::reset();
    _F = -1;
  }

  // ---- used by OverSampledDFTSynthesisBank for the fused path
  bool all_own_banks(unsigned& m, unsigned& r, unsigned& dct) const {
    if (_channelList.empty()) return false;
    const OverSampledDFTAnalysisBank* a0 = 0;
    for (_ChannelList::const_iterator it = _channelList.begin(); it != _channelList.end(); ++it) {
      const OverSampledDFTAnalysisBank* a = dynamic_cast<const OverSampledDFTAnalysisBank*>(it->get());
      if (!a || a->frameX() >= 0 || a->_M != _fftLen) return false;
      if (!a0) a0 = a;
      if (a->_m != a0->_m || a->_r != a0->_r || a->_dct != a0->_dct) return false;
    }
    m = a0->_m; r = a0->_r; dct = a0->_dct;
    return true;
  }
  const std::vector<double>& analysis_prototype() const {
    return dynamic_cast<const OverSampledDFTAnalysisBank*>(_channelList.front().get())->_prototype;
  }
  void interleaved_pcm(std::vector<float>& pcm, long& T) {
    const unsigned C = chanN();
    std::vector<std::vector<float> > xs(C);
    unsigned c = 0;
    T = 0;
    for (_ChannelIterator it = _channelList.begin(); it != _channelList.end(); ++it, ++c) {
      dynamic_cast<OverSampledDFTAnalysisBank*>(it->get())->pull_source(xs[c]);
      if ((long)xs[c].size() > T) T = (long)xs[c].size();
    }
    pcm.assign((size_t)T * C, 0.f);
    for (c = 0; c < C; c++) for (size_t t = 0; t < xs[c].size(); t++) pcm[t * C + c] = xs[c][t];
  }
  btkb200_plan* plan() { return need_plan(); }
  // beamformer output multiplied by the Zelinski gain for every frame of the utterance (ZelinskiPostFilter::next)
  void post_filter(double alpha, int type, int minFrames, std::vector<float>& Y, std::vector<float>& W, int& F) {
    if (_F < 0) evaluate();
    const unsigned B = _fftLen / 2 + 1;
    Y.assign((size_t)_F * B * 2 + 2, 0.f);
    W.assign((size_t)_F * B + 1, 0.f);
    if (_F > 0) check(btkb200_beamform_zelinski(_plan.get(), &_snapshots[0], _F, alpha, type, minFrames, &Y[0], &W[0]), _plan.get());
    F = _F;
  }
  virtual void require_weights() {
    if (!need_plan() || !_plan.has_weights()) throw j_error("call calcArrayManifoldVectorsX() once\n");   // :1140-1143
  }

 protected:
  typedef std::list<VectorComplexFeatureStreamPtr> _ChannelList;
  typedef _ChannelList::iterator _ChannelIterator;

  static void fill(btk_vector_complex* v, unsigned n) {
    if (v->size != n) { delete[] v->data; v->data = new double[2 * (size_t)n](); v->size = n; }
  }
  btkb200_plan* need_plan() {
    const unsigned C = chanN();
    if (C == 0) throw j_error("No channel is set");
    if (!_plan || _plan.C() != C) {
      unsigned m = 1, r = 0, dct = 0;
      if (all_own_banks(m, r, dct)) _plan.create(_fftLen, m, r, dct, C, &analysis_prototype()[0], 0, 1);
      else _plan.create(_fftLen, 1, 0, 0, C, 0, 0, 1);
    }
    return _plan.get();
  }
  void evaluate() {
    require_weights();
    const unsigned C = chanN(), B = _fftLen / 2 + 1;
    unsigned m, r, dct;
    long F = 0;
    if (all_own_banks(m, r, dct)) {       // one multichannel analysis launch
      std::vector<float> pcm; long T;
      interleaved_pcm(pcm, T);
      F = btkb200_analysis_frames(_plan.get(), T);
      _snapshots.assign((size_t)F * B * C * 2, 0.f);
      float dummy = 0.f;
      check(btkb200_analysis(_plan.get(), pcm.empty() ? &dummy : &pcm[0], T, &_snapshots[0], &F), _plan.get());
    } else {                              // foreign upstream nodes: pull their frames one by one (beamformer.cc:1152-1156)
      std::vector<std::vector<double> > per(C);
      unsigned c = 0;
      F = -1;
      for (_ChannelIterator it = _channelList.begin(); it != _channelList.end(); ++it, ++c) {
        long n = 0;
        for (;;) {
          const btk_vector_complex* x;
          try { x = (*it)->next(); } catch (jiterator_error&) { break; }
          for (unsigned s = 0; s < B; s++) { per[c].push_back(x->data[2 * s * x->stride]); per[c].push_back(x->data[2 * s * x->stride + 1]); }
          n++;
        }
        if (F < 0 || n < F) F = n;
      }
      _snapshots.assign((size_t)F * B * C * 2, 0.f);
      for (c = 0; c < C; c++)
        for (long f = 0; f < F; f++)
          for (unsigned s = 0; s < B; s++) {
            _snapshots[(((size_t)f * B + s) * C + c) * 2] = (float)per[c][((size_t)f * B + s) * 2];
            _snapshots[(((size_t)f * B + s) * C + c) * 2 + 1] = (float)per[c][((size_t)f * B + s) * 2 + 1];
          }
    }
    _Y.assign((size_t)F * B * 2 + 2, 0.f);
    if (F > 0) check(btkb200_beamform(_plan.get(), &_snapshots[0], F, &_Y[0]), _plan.get());
    _F = (int)F;
  }

  unsigned _fftLen, _fftLen2;
  _ChannelList _channelList;
  PlanHandle _plan;
  std::vector<float> _snapshots, _Y;
  int _F;
  btk_vector_complex* _snap;
  btk_vector_complex* _wvec;
};

class SubbandDS : public SubbandBeamformer {
 public:
  SubbandDS(unsigned fftLen = 512, bool halfBandShift = false, const String& nm = "SubbandDS")
      : SubbandBeamformer(fftLen, halfBandShift, nm) {}
  // beamformer.cc:1087-1091 -> beamformerWeights::calcMainlobe (:531-594)
  virtual void calcArrayManifoldVectors(double sampleRate, const btk_vector* delays) {
    if (delays->size != chanN())
      throw jdimension_error("Number of delays does not match number of channels (%d vs. %d).\n", (int)delays->size, (int)chanN());
    std::vector<double> d(delays->size);
    for (size_t i = 0; i < d.size(); i++) d[i] = delays->data[i * delays->stride];
    check(btkb200_set_ds_weights(need_plan(), sampleRate, &d[0], (unsigned)d.size()), _plan.get());
    _F = -1;
  }
  virtual const btk_vector_complex* getWeights(unsigned fbinX) {
    const unsigned C = chanN(), B = _fftLen / 2 + 1;
    std::vector<double> w((size_t)B * C * 2);
    check(btkb200_get_weights(need_plan(), &w[0]), _plan.get());
    fill(_wvec, C);
    memcpy(_wvec->data, &w[(size_t)fbinX * C * 2], sizeof(double) * 2 * C);
    return _wvec;
  }
};

class SubbandMVDR : public SubbandDS {
 public:
  SubbandMVDR(unsigned fftLen = 512, bool halfBandShift = false, const String& nm = "SubbandMVDR")
      : SubbandDS(fftLen, halfBandShift, nm) {}
  bool setNoiseSpatialSpectralMatrix(unsigned fbinX, btk_matrix_complex* Rnn) {   // beamformer.cc:2454-2477
    const unsigned C = chanN();
    if (Rnn->size1 != C || Rnn->size2 != C) return false;
    std::vector<double> R((size_t)C * C * 2);
    for (unsigned a = 0; a < C; a++) memcpy(&R[(size_t)a * C * 2], Rnn->data + (size_t)a * Rnn->tda * 2, sizeof(double) * 2 * C);
    return btkb200_set_covariance(need_plan(), fbinX, &R[0], C, C) == BTKB200_OK;
  }
  bool setDiffuseNoiseModel(const btk_matrix* micPositions, double sampleRate, double sspeed = 343740.0) {   // :2486-2553
    const unsigned C = chanN();
    if (micPositions->size1 != C || micPositions->size2 < 3) return false;
    std::vector<double> mp((size_t)C * 3);
    for (unsigned c = 0; c < C; c++) for (unsigned k = 0; k < 3; k++) mp[c * 3 + k] = micPositions->data[c * micPositions->tda + k];
    return btkb200_set_diffuse_noise_model(need_plan(), &mp[0], C, sampleRate, sspeed) == BTKB200_OK;
  }
  void setAllLevelsOfDiagonalLoading(float diagonalWeight) { check(btkb200_diag_load(need_plan(), diagonalWeight), _plan.get()); }
  void setLevelOfDiagonalLoading(unsigned fbinX, float diagonalWeight) { check(btkb200_diag_load_bin(need_plan(), fbinX, diagonalWeight), _plan.get()); }
  void divideAllNonDiagonalElements(float myu) { check(btkb200_divide_nondiagonal(need_plan(), myu), _plan.get()); }
  bool calcMVDRWeights(double sampleRate, double dThreshold = 1.0E-8, bool /*calcInverseMatrix*/ = true) {   // :2392-2446
    int nfb = 0;
    check(btkb200_solve_mvdr(need_plan(), sampleRate, dThreshold, &nfb), _plan.get());
    _F = -1;
    _mvdr_ready = true;
    return true;
  }
  // SubbandMVDR::next (beamformer.cc:2587-2594): the manifold first, then the MVDR weights; delay-and-sum output is never
  // served in place of MVDR output
  virtual void require_weights() {
    SubbandBeamformer::require_weights();
    if (!_mvdr_ready) throw j_error("call calcMVDRWeights() once\n");
  }

 private:
  bool _mvdr_ready = false;

 public:
  const btk_vector_complex* getMVDRWeights(unsigned fbinX) { return getWeights(fbinX); }
};
typedef std::shared_ptr<SubbandDS> SubbandDSPtr;
typedef std::shared_ptr<SubbandMVDR> SubbandMVDRPtr;

// SubbandGSC with FIXED active weights (beamformer/beamformer.h:186-210, beamformer.cc:1296-1447); the adaptive
// subclasses (SubbandGSCRLS, ...) are out of scope.  The effective weights wq - B wa are installed before evaluation.
class SubbandGSC : public SubbandDS {
 public:
  SubbandGSC(unsigned fftLen = 512, bool halfBandShift = false, const String& nm = "SubbandGSC")
      : SubbandDS(fftLen, halfBandShift, nm), _normalizeWeight(false), _dirty(false) {
    _bm = new btk_matrix_complex(); _bm->size1 = _bm->size2 = _bm->tda = 0; _bm->data = 0; _bm->block = 0; _bm->owner = 0;
  }
  ~SubbandGSC() { delete[] _bm->data; delete _bm; }
  void normalizeWeight(bool flag) { _normalizeWeight = flag; _dirty = true; _F = -1; }
  void calcGSCWeights(double sampleRate, const btk_vector* delaysT) {          // :1373-1377
    if (delaysT->size != chanN())
      throw jdimension_error("Number of delays does not match number of channels (%d vs. %d).\n", (int)delaysT->size, (int)chanN());
    std::vector<double> d(delaysT->size);
    for (size_t i = 0; i < d.size(); i++) d[i] = delaysT->data[i * delaysT->stride];
    check(btkb200_gsc_calc_weights(need_plan(), sampleRate, &d[0], (unsigned)d.size()), _plan.get());
    _dirty = true; _F = -1;
  }
  void setActiveWeights_f(unsigned fbinX, const btk_vector* packedWeight) {   // :1425-1433
    std::vector<double> w(packedWeight->size);
    for (size_t i = 0; i < w.size(); i++) w[i] = packedWeight->data[i * packedWeight->stride];
    if (fbinX <= _fftLen2) check(btkb200_gsc_set_active_weights(need_plan(), fbinX, &w[0], (unsigned)w.size()), _plan.get());
    _dirty = true; _F = -1;
  }
  void zeroActiveWeights() { check(btkb200_gsc_zero_active_weights(need_plan()), _plan.get()); _dirty = true; _F = -1; }   // :1435-1447
  btk_matrix_complex* getBlockingMatrix(unsigned /*srcX*/, unsigned fbinX) {
    const unsigned C = chanN();
    if (_bm->size1 != C) { delete[] _bm->data; _bm->data = new double[2 * (size_t)C * (C - 1)](); _bm->size1 = C; _bm->size2 = _bm->tda = C - 1; }
    check(btkb200_gsc_get_blocking_matrix(need_plan(), fbinX, _bm->data), _plan.get());
    return _bm;
  }
  virtual const btk_vector_complex* next(int frameX = -5) {
    if (_dirty) { check(btkb200_gsc_apply(need_plan(), _normalizeWeight ? 1 : 0), _plan.get()); _dirty = false; _F = -1; }
    return SubbandDS::next(frameX);
  }
  // the fused synthesis path asks for the weights through plan(): make sure they are current
  void commit() { if (_dirty) { check(btkb200_gsc_apply(need_plan(), _normalizeWeight ? 1 : 0), _plan.get()); _dirty = false; } }

 protected:
  bool _normalizeWeight, _dirty;
  btk_matrix_complex* _bm;
};
typedef std::shared_ptr<SubbandGSC> SubbandGSCPtr;

// ZelinskiPostFilter (postfilter/postfilter.h:95-126, postfilter.cc:340-500): the node both shipped drivers put between the
// beamformer and the synthesis bank.  `output` must be the SubbandDS / SubbandMVDR also passed to setBeamformer().
enum PostfilterType { TYPE_ZELINSKI1_REAL = 0x01, TYPE_ZELINSKI1_ABS = 0x02, TYPE_APAB = 0x04, TYPE_ZELINSKI2 = 0x08, NO_USE_POST_FILTER = 0x00 };
class ZelinskiPostFilter : public VectorComplexFeatureStream {
 public:
  ZelinskiPostFilter(VectorComplexFeatureStreamPtr& output, unsigned fftLen, double alpha = 0.6, int type = 2, int minFrames = 0,
                     const String& nm = "ZelinskPostFilter")
      : VectorComplexFeatureStream(fftLen, nm), _samp(output), _alpha(alpha), _type(type), _minFrames(minFrames), _F(-1) {
    if (output->size() != fftLen)      // postfilter.cc:355-358
      throw jdimension_error("Input block length (%d) != fftLen (%d)\n", (int)output->size(), (int)fftLen);
    _wp1 = new btk_vector_complex(); _wp1->size = fftLen; _wp1->stride = 1; _wp1->data = new double[2 * (size_t)fftLen](); _wp1->block = 0; _wp1->owner = 1;
  }
  ~ZelinskiPostFilter() { delete[] _wp1->data; delete _wp1; }
  void setBeamformer(SubbandDSPtr& beamformer) { _bf = beamformer; }
  const btk_vector_complex* getPostFilterWeights() { return (_F < 0 || _frameX < 0) ? 0 : _wp1; }
  virtual const btk_vector_complex* next(int frameX = -5) {
    if (frameX == _frameX && _frameX >= 0) return _vector;
    if (_F < 0) evaluate();
    const int t = _frameX + 1;
    if (t >= _F) { _endOfSamples = true; throw jiterator_error("end of samples!"); }
    const unsigned M = size(), B = M / 2 + 1;
    const float* half = &_Y[(size_t)t * B * 2];
    const float* wh = &_W[(size_t)t * B];
    for (unsigned s = 0; s < B; s++) {
      _vector->data[2 * s] = half[2 * s]; _vector->data[2 * s + 1] = half[2 * s + 1];
      _wp1->data[2 * s] = wh[s]; _wp1->data[2 * s + 1] = 0.0;
    }
    for (unsigned s = 1; s < M / 2; s++) {                     // postfilter.cc:185-186, 213-216
      _vector->data[2 * (M - s)] = half[2 * s]; _vector->data[2 * (M - s) + 1] = -half[2 * s + 1];
      _wp1->data[2 * (M - s)] = wh[s]; _wp1->data[2 * (M - s) + 1] = 0.0;
    }
    _increment();
    return _vector;
  }
  virtual void reset() { _samp->reset(); VectorComplexFeatureStream::reset(); _F = -1; }

 private:
  void evaluate() {
    if (!_bf) throw j_error("set beamformer's weights \n");   // postfilter.cc:449-452
    if (static_cast<VectorComplexFeatureStream*>(_bf.get()) != _samp.get())
      throw j_error("the B200 post-filter expects its input stream to be the beamformer given to setBeamformer()");
    _bf->post_filter(_alpha, _type, _minFrames, _Y, _W, _F);
  }
  VectorComplexFeatureStreamPtr _samp;
  SubbandDSPtr _bf;
  double _alpha;
  int _type, _minFrames, _F;
  std::vector<float> _Y, _W;
  btk_vector_complex* _wp1;
};
typedef std::shared_ptr<ZelinskiPostFilter> ZelinskiPostFilterPtr;

// ---- OverSampledDFTSynthesisBank ---------------------------------------------------------------------------------
class OverSampledDFTSynthesisBank : public OverSampledDFTFilterBank, public VectorFloatFeatureStream {
 public:
  OverSampledDFTSynthesisBank(VectorComplexFeatureStreamPtr& samp, btk_vector* prototype, unsigned M, unsigned m, unsigned r = 0,
                              unsigned delayCompensationType = 0, int gainFactor = 1,
                              const String& nm = "OverSampledDFTSynthesisBank")
      : OverSampledDFTFilterBank(prototype, M, m, r, delayCompensationType), VectorFloatFeatureStream(M >> r, nm), _samp(samp),
        _gain(gainFactor), _noStreamFeature(false), _nout(-1), _fused(false) {}
  // push-style use: inputSourceVector() + next()  (modulated.h:348-352, modulated.cc:581-593)
  OverSampledDFTSynthesisBank(btk_vector* prototype, unsigned M, unsigned m, unsigned r = 0, unsigned delayCompensationType = 0,
                              int gainFactor = 1, const String& nm = "OverSampledDFTSynthesisBank")
      : OverSampledDFTFilterBank(prototype, M, m, r, delayCompensationType), VectorFloatFeatureStream(M >> r, nm), _gain(gainFactor),
        _noStreamFeature(true), _nout(-1), _fused(false) {}

  void doNotUseStreamFeature(bool flag = true) { _noStreamFeature = flag; }
  void inputSourceVector(const btk_vector_complex* block) {
    const unsigned B = _M / 2 + 1;
    for (unsigned s = 0; s < B; s++) { _pushed.push_back((float)block->data[2 * s * block->stride]); _pushed.push_back((float)block->data[2 * s * block->stride + 1]); }
  }
  bool fused() const { return _fused; }

  virtual const btk_vector_float* next(int frameX = -5) {
    if (frameX == _frameX && _frameX >= 0) return _vector;
    if (_noStreamFeature) return next_pushed();
    if (_nout < 0) evaluate();
    const int t = _frameX + 1;
    if (t >= _nout) { _endOfSamples = true; throw jiterator_error("end of samples!"); }
    memcpy(_vector->data, &_out[(size_t)t * _D], sizeof(float) * _D);
    _increment();
    return _vector;
  }
  virtual void reset() {
    if (_samp) _samp->reset();
    VectorFloatFeatureStream::reset();
    _nout = -1;
    _pushed.clear();
  }

 private:
  void need_plan(unsigned C, const double* h) {
    if (!_plan || _plan.C() != C) _plan.create(_M, _m, _r, _dct, C, h, &_prototype[0], _gain);
  }
  void evaluate() {
    SubbandBeamformer* bf = dynamic_cast<SubbandBeamformer*>(_samp.get());
    unsigned m, r, dct;
    if (bf && bf->frameX() < 0 && bf->fftLen() == _M && bf->all_own_banks(m, r, dct) && m == _m && r == _r && dct == _dct) {
      // ---- analysis -> weights -> synthesis as ONE kernel
      if (SubbandGSC* gsc = dynamic_cast<SubbandGSC*>(bf)) gsc->commit();   // install wq - B wa before the weights are read
      bf->require_weights();
      const unsigned C = bf->chanN(), B = _M / 2 + 1;
      need_plan(C, &bf->analysis_prototype()[0]);
      std::vector<double> w((size_t)B * C * 2);
      check(btkb200_get_weights(bf->plan(), &w[0]), bf->plan());
      check(btkb200_set_weights(_plan.get(), &w[0]), _plan.get());
      std::vector<float> pcm; long T;
      bf->interleaved_pcm(pcm, T);
      const long nblk = btkb200_chain_frames(_plan.get(), T);
      _out.assign((size_t)nblk * _D + 1, 0.f);
      float dummy = 0.f;
      check(btkb200_chain(_plan.get(), pcm.empty() ? &dummy : &pcm[0], T, &_out[0]), _plan.get());
      _nout = (int)nblk;
      _fused = true;
      return;
    }
    // ---- foreign upstream: pull its frames (jiterator_error = end of stream, modulated.cc:639-642)
    const unsigned B = _M / 2 + 1;
    std::vector<float> Y;
    long F = 0;
    for (;;) {
      const btk_vector_complex* y;
      try { y = _samp->next(); } catch (jiterator_error&) { break; }
      for (unsigned s = 0; s < B; s++) { Y.push_back((float)y->data[2 * s * y->stride]); Y.push_back((float)y->data[2 * s * y->stride + 1]); }
      F++;
    }
    run_synthesis(Y, F);
    _fused = false;
  }
  void run_synthesis(std::vector<float>& Y, long F) {
    need_plan(1, 0);
    long nout = btkb200_synthesis_frames(_plan.get(), F);
    _out.assign((size_t)nout * _D + 1, 0.f);
    if (Y.empty()) Y.push_back(0.f);
    check(btkb200_synthesis(_plan.get(), &Y[0], F, &_out[0], &nout), _plan.get());
    _nout = (int)nout;
  }
  // one output per call from the frames pushed so far; only the newest output is needed, so a short window of
  // trailing frames is re-run (exact once the window is longer than the filter memory; from the stream start
  // the whole history is used, which reproduces the priming behaviour of modulated.cc:626-664)
  const btk_vector_float* next_pushed() {
    const unsigned B = _M / 2 + 1;
    const long F = (long)(_pushed.size() / (2 * B));
    need_plan(1, 0);
    btkb200_info inf; btkb200_plan_info(_plan.get(), &inf);
    const long pd = inf.pd_synthesis;
    const long j = _frameX + 1;                       // output frame wanted: needs frames up to j + pd
    if (F < j + pd + 1) throw jiterator_error("end of samples!");
    long lead = (long)_m * _R + _R + pd;              // frames before j+pd that influence out_j, with margin
    long f0 = j + pd + 1 - lead - pd - 1;
    if (f0 < 0) f0 = 0;
    std::vector<float> Y(_pushed.begin() + (size_t)f0 * 2 * B, _pushed.begin() + (size_t)(j + pd + 1) * 2 * B);
    if (f0 == 0) { run_synthesis(Y, j + pd + 1); memcpy(_vector->data, &_out[(size_t)j * _D], sizeof(float) * _D); }
    else { run_synthesis(Y, j + pd + 1 - f0); memcpy(_vector->data, &_out[(size_t)(_nout - 1) * _D], sizeof(float) * _D); }
    _nout = -1;
    _increment();
    return _vector;
  }

  VectorComplexFeatureStreamPtr _samp;
  int _gain;
  bool _noStreamFeature;
  PlanHandle _plan;
  std::vector<float> _out, _pushed;
  int _nout;
  bool _fused;
};
typedef std::shared_ptr<OverSampledDFTSynthesisBank> OverSampledDFTSynthesisBankPtr;

}  // namespace btkb200
#endif  // BTKB200_STREAMS_H
