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
// Two build modes.
//   * stand-alone (default): GSL and the reference tree are not needed.  The header defines containers with the SAME
//     field layout as gsl_vector_float / gsl_vector_complex / gsl_vector / gsl_matrix(_complex) under the names
//     btk_vector_*, and its own FeatureStream / j_error family / SnapShotArray with the reference's members.
//   * BTKB200_WITH_BTK: compiled INSIDE the reference tree (include path = btk/ and GSL).  Nothing is re-declared: the
//     nodes derive from the reference's own FeatureStream<> (stream/stream.h:35-107), are held by its refcountable_ptr
//     (common/refcount.h:186-285), throw its j_error family (common/jexception.h:41-173), and the beamformer derives from the
//     reference's ::SubbandDS (beamformer/beamformer.h:159-182), keeping its SnapShotArray and beamformerWeights objects
//     current -- so a reference node (ZelinskiPostFilter(VectorComplexFeatureStreamPtr&, ...) + setBeamformer(SubbandDSPtr&),
//     OverSampledDFTSynthesisBank, ...) takes a B200 node and the reverse.  The test tier builds this mode against the
//     reference's headers and runs mixed chains (tests/host/test_mixed_chain.cc); INTEGRATION.md shows the build line
//     inside the reference tree.
#ifndef BTKB200_STREAMS_H
#define BTKB200_STREAMS_H

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <exception>
#include <list>
#include <memory>
#include <string>
#include <vector>

#include "../../include/btkb200.h"

#if defined(BTKB200_WITH_BTK)
#include <gsl/gsl_matrix.h>
#include <gsl/gsl_vector.h>
#include "stream/stream.h"
#include "beamformer/beamformer.h"
typedef gsl_vector_char btk_vector_char;
typedef gsl_vector_float btk_vector_float;
typedef gsl_vector_complex btk_vector_complex;
typedef gsl_vector btk_vector;
typedef gsl_matrix btk_matrix;
typedef gsl_matrix_complex btk_matrix_complex;
#else
struct btk_vector_char { size_t size, stride; char* data; void* block; int owner; };
struct btk_vector_float { size_t size, stride; float* data; void* block; int owner; };
struct btk_vector_complex { size_t size, stride; double* data; void* block; int owner; };   // (re, im) pairs
struct btk_vector { size_t size, stride; double* data; void* block; int owner; };
struct btk_matrix { size_t size1, size2, tda; double* data; void* block; int owner; };
struct btk_matrix_complex { size_t size1, size2, tda; double* data; void* block; int owner; };
#endif

namespace btkb200 {

#if defined(BTKB200_WITH_BTK)
// ---- everything below comes from the reference's own headers ---------------------------------------------------
typedef ::String String;
using ::j_error;
using ::jiterator_error;
using ::jconsistency_error;
using ::jdimension_error;
using ::jio_error;
using ::VectorCharFeatureStream;
using ::VectorCharFeatureStreamPtr;
using ::VectorFloatFeatureStream;
using ::VectorComplexFeatureStream;
using ::VectorFloatFeatureStreamPtr;
using ::VectorComplexFeatureStreamPtr;
using ::SnapShotArray;
using ::SnapShotArrayPtr;
using ::beamformerWeights;
template <class T> struct node_ptr { typedef refcountable_ptr<T> type; };
template <class T> inline T* pget(const refcountable_ptr<T>& p) { return p.isNull() ? 0 : p.operator->(); }
template <class T> inline T* pget(const refcount_ptr<T>& p) { return p.isNull() ? 0 : p.operator->(); }
#else
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
BTKB200_DEFINE_ERROR(jio_error, JIO)
#undef BTKB200_DEFINE_ERROR

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

class VectorCharFeatureStream : public FeatureStream<btk_vector_char, char> {
 protected:
  VectorCharFeatureStream(unsigned sz, const String& nm) : FeatureStream<btk_vector_char, char>(sz, nm, 1) {}
};
typedef std::shared_ptr<VectorCharFeatureStream> VectorCharFeatureStreamPtr;
class VectorFloatFeatureStream : public FeatureStream<btk_vector_float, float> {
 protected:
  VectorFloatFeatureStream(unsigned sz, const String& nm) : FeatureStream<btk_vector_float, float>(sz, nm, 1) {}
};
class VectorComplexFeatureStream : public FeatureStream<btk_vector_complex, double> {
 protected:
  VectorComplexFeatureStream(unsigned sz, const String& nm) : FeatureStream<btk_vector_complex, double>(sz, nm, 2) {}
};
template <class T> struct node_ptr { typedef std::shared_ptr<T> type; };
template <class T> inline T* pget(const std::shared_ptr<T>& p) { return p.get(); }
typedef std::shared_ptr<VectorFloatFeatureStream> VectorFloatFeatureStreamPtr;
typedef std::shared_ptr<VectorComplexFeatureStream> VectorComplexFeatureStreamPtr;
#endif

// a vector of n complex values with the gsl_vector_complex field layout
inline btk_vector_complex* new_cvec(size_t n) {
  btk_vector_complex* v = new btk_vector_complex();
  v->size = n; v->stride = 1; v->data = n ? new double[2 * n]() : 0; v->block = 0; v->owner = 0;
  return v;
}
inline void free_cvec(btk_vector_complex* v) { if (v) { delete[] v->data; delete v; } }

// status of the C ABI -> the exception the reference throws at that point (include/btkb200.h)
inline void check(int rc, const btkb200_plan* plan) {
  if (rc == BTKB200_OK) return;
  const char* msg = btkb200_last_error(plan);
  if (rc == BTKB200_EINVAL) throw jdimension_error("%s", msg);
  throw j_error("%s", msg);
}

// ---- G1: delay helpers of the reference's drivers ----------------------------------------------------
// calcDelaysPolar2 (src/superdirectiveBeamformer.cc:118-137): far-field delays in seconds, micPos [C][3] in mm
inline void calcDelaysPolar2(float azimuth, float elevation, const btk_matrix* micPos, btk_vector* delays) {
  std::vector<double> mp(micPos->size1 * 3), d(micPos->size1);
  for (size_t c = 0; c < micPos->size1; c++) for (int k = 0; k < 3; k++) mp[c * 3 + k] = micPos->data[c * micPos->tda + k];
  check(btkb200_calc_delays_polar(azimuth, elevation, &mp[0], (unsigned)micPos->size1, &d[0]), 0);
  for (size_t c = 0; c < d.size(); c++) delays->data[c * delays->stride] = d[c];
}
// calcAllDelays (beamformer/beamformer.cc:1214-1231), the ignored source position included
inline void calcAllDelays(double x, double y, double z, const btk_matrix* mpos, btk_vector* delays) {
  std::vector<double> mp(mpos->size1 * 3), d(mpos->size1);
  for (size_t c = 0; c < mpos->size1; c++) for (int k = 0; k < 3; k++) mp[c * 3 + k] = mpos->data[c * mpos->tda + k];
  check(btkb200_calc_all_delays(x, y, z, &mp[0], (unsigned)mpos->size1, &d[0]), 0);
  for (size_t c = 0; c < d.size(); c++) delays->data[c * delays->stride] = d[c];
}

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

// ---- raw-PCM ingest nodes -------------------------------------------------------------------------------
// byte block source over memory (what the Mark-III driver node delivers: packed big-endian 24-bit samples)
class MemoryCharFeature : public VectorCharFeatureStream {
 public:
  MemoryCharFeature(const char* data, size_t n, unsigned blockLen, const String& nm = "MemoryChar")
      : VectorCharFeatureStream(blockLen, nm), _d(data, data + n), _cur(0) {}
  virtual const btk_vector_char* next(int frameX = -5) {
    if (frameX == _frameX) return _vector;
    if (_cur + _size > _d.size()) { _endOfSamples = true; throw jiterator_error("end of samples!"); }
    memcpy(_vector->data, &_d[_cur], _size);
    _cur += _size;
    _increment();
    return _vector;
  }
  virtual void reset() { VectorCharFeatureStream::reset(); _cur = 0; }
 private:
  std::vector<char> _d;
  size_t _cur;
};

// Conversion24bit2Float (feature/feature.h:148-158, feature.cc:190-217): 3 bytes, most significant first, -> float.  The byte
// stream is drained once and widened on the device in one call (btkb200_convert_pcm, bit-exact), then served in blocks.
class Conversion24bit2Float : public VectorFloatFeatureStream {
 public:
  Conversion24bit2Float(VectorCharFeatureStreamPtr& src, const String& nm = "Conversion from 24 bit integer to Float")
      : VectorFloatFeatureStream(src->size() / 3, nm), _src(src), _ready(false) {}
  virtual void reset() { _src->reset(); VectorFloatFeatureStream::reset(); _ready = false; }
  virtual const btk_vector_float* next(int frameX = -5) {
    if (frameX == _frameX && _frameX >= 0) return _vector;
    if (frameX >= 0 && frameX - 1 != _frameX)
      throw jconsistency_error("Problem in Feature %s: %d != %d\n", name().c_str(), frameX - 1, _frameX);
    if (!_ready) {
      std::vector<char> raw;
      for (;;) {
        const btk_vector_char* b;
        try { b = _src->next(); } catch (jiterator_error&) { break; }
        raw.insert(raw.end(), b->data, b->data + b->size);
      }
      const long n = (long)(raw.size() / 3);
      _all.assign((size_t)n + 1, 0.f);
      if (n > 0) {
        PlanHandle plan;
        plan.create(64, 1, 0, 0, 1, 0, 0, 1);
        check(btkb200_convert_pcm(plan.get(), &raw[0], BTKB200_PCM_S24BE, n, &_all[0]), plan.get());
      }
      _all.resize((size_t)n);
      _ready = true;
    }
    const size_t t = (size_t)(_frameX + 1);
    if ((t + 1) * _size > _all.size()) { _endOfSamples = true; throw jiterator_error("end of samples!"); }
    memcpy(_vector->data, &_all[t * _size], sizeof(float) * _size);
    _increment();
    return _vector;
  }
 private:
  VectorCharFeatureStreamPtr _src;
  std::vector<float> _all;
  bool _ready;
};

// ChannelExtractionFeature (feature/feature.h:1823-1841, feature.cc:3885-3900): out[i] = in[i chN + chX]
class ChannelExtractionFeature : public VectorFloatFeatureStream {
 public:
  ChannelExtractionFeature(const VectorFloatFeatureStreamPtr& src, unsigned chX = 0, unsigned chN = 1, const String& nm = "ChannelExtraction")
      : VectorFloatFeatureStream(src->size() / chN, nm), _src(src), _chX(chX), _chN(chN) {
    if (chX >= chN || src->size() % chN != 0) throw jdimension_error("channel %d of %d out of blocks of %d", (int)chX, (int)chN, (int)src->size());
  }
  virtual void reset() { _src->reset(); VectorFloatFeatureStream::reset(); }
  virtual const btk_vector_float* next(int frameX = -5) {
    if (frameX == _frameX && _frameX >= 0) return _vector;
    if (frameX >= 0 && frameX - 1 != _frameX)
      throw jconsistency_error("Problem in Feature %s: %d != %d\n", name().c_str(), frameX - 1, _frameX);
    _increment();
    const btk_vector_float* all = _src->next(_frameX);
    for (unsigned i = 0; i < _size; i++) _vector->data[i] = all->data[(i * _chN + _chX) * all->stride];
    return _vector;
  }
 private:
  VectorFloatFeatureStreamPtr _src;
  unsigned _chX, _chN;
};

// IterativeSampleFeature (feature/feature.h:301-335, feature.cc:803-896): one node per channel of ONE interleaved file.  The
// nodes share the file image (process-wide, like the reference's static members) and are pulled in lock step; the node of
// channel firstChanX refills a 30-s buffer when its block counter wraps, so the stream is a whole number of buffers long,
// zero padded, and ends at the first wrap after a short read (:880-892).  read() decodes RIFF/WAVE 16-bit PCM (or a
// headerless 16-bit file with the given channel count) and widens it on the device (bit-exact; SFC_SET_NORM_FLOAT is off in
// the reference, :849).  An analysis bank fed by such a node takes the channel's complete stream from whole_stream().
class IterativeSampleFeature : public VectorFloatFeatureStream {
 public:
  IterativeSampleFeature(unsigned chX, unsigned blockLen = 320, unsigned firstChanX = 0, const String& nm = "Iterative Sample")
      : VectorFloatFeatureStream(blockLen, nm), _blockLen(blockLen), _chanX(chX), _firstChanX(firstChanX), _cur(0), _last(false), _cto(-1) {}
  void read(const String& fileName, int /*format*/ = 0, int samplerate = 44100, int chN = 1, int cfrom = 0, int cto = -1) {
    if (_chanX != _firstChanX) return;
    FILE* fp = fopen(fileName.c_str(), "rb");
    if (!fp) throw jio_error("Could not open file %s.", fileName.c_str());
    std::vector<unsigned char> bytes;
    unsigned char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, fp)) > 0) bytes.insert(bytes.end(), buf, buf + n);
    fclose(fp);
    size_t off = 0, len = bytes.size();
    if (len >= 12 && !memcmp(&bytes[0], "RIFF", 4) && !memcmp(&bytes[8], "WAVE", 4)) {
      size_t p = 12;
      bool have_fmt = false;
      off = len = 0;
      while (p + 8 <= bytes.size()) {
        const unsigned sz = bytes[p + 4] | (bytes[p + 5] << 8) | (bytes[p + 6] << 16) | ((unsigned)bytes[p + 7] << 24);
        if (!memcmp(&bytes[p], "fmt ", 4) && sz >= 16) {
          const unsigned fmt = bytes[p + 8] | (bytes[p + 9] << 8), bits = bytes[p + 22] | (bytes[p + 23] << 8);
          chN = bytes[p + 10] | (bytes[p + 11] << 8);
          samplerate = (int)(bytes[p + 12] | (bytes[p + 13] << 8) | (bytes[p + 14] << 16) | ((unsigned)bytes[p + 15] << 24));
          if (fmt != 1 || bits != 16) throw jio_error("Could not open file %s: only 16-bit PCM WAVE files are decoded", fileName.c_str());
          have_fmt = true;
        } else if (!memcmp(&bytes[p], "data", 4)) {
          off = p + 8; len = sz < bytes.size() - off ? sz : bytes.size() - off;
          break;
        }
        p += 8 + sz + (sz & 1);
      }
      if (!have_fmt || off == 0) throw jio_error("sndfile error: %s.", "no fmt / data chunk");
    }
    readRaw(reinterpret_cast<const short*>(&bytes[off]), (long)(len / 2), samplerate, chN, cfrom, cto);
  }
  // the same with the interleaved 16-bit samples already in memory (n = number of int16 values)
  void readRaw(const short* raw, long n, int samplerate, int chN, int cfrom = 0, int cto = -1) {
    if (_chanX != _firstChanX) return;
    if (cto > 0 && cto < cfrom) throw jconsistency_error("Segment cannot start at %d and end at %d", cfrom, cto);
    Shared& S = shared();
    n = n / chN * chN;
    S.pcm.assign((size_t)n + 1, 0.f);
    if (n > 0) {
      PlanHandle plan;
      plan.create(64, 1, 0, 0, 1, 0, 0, 1);
      check(btkb200_convert_pcm(plan.get(), raw, BTKB200_PCM_S16, n, &S.pcm[0]), plan.get());
    }
    S.pcm.resize((size_t)n);
    S.chN = chN; S.samplerate = samplerate; S.pos = cfrom; S.cfrom = cfrom;
    S.blockN = S.interval * (unsigned)samplerate / _blockLen + 1;
    S.sampleN = S.blockN * _blockLen;
    S.buf.assign((size_t)S.sampleN * chN, 0.f);
    _cto = cto - cfrom;
  }
  unsigned samplesN() const { return shared().ttl; }
  void changeFirstChannelID(unsigned firstChanX) { _firstChanX = firstChanX; }
  virtual void reset() { shared().ttl = 0; _cur = 0; _last = false; VectorFloatFeatureStream::reset(); }
  virtual const btk_vector_float* next(int frameX = -5) {
    if (frameX == _frameX && _frameX >= 0) return _vector;
    Shared& S = shared();
    if (S.buf.empty()) throw jio_error("no file has been read");
    const unsigned currentFrame = _cur % S.blockN;
    if (_chanX == _firstChanX && currentFrame == 0) {
      if (_last || (_cto > 0 && (long)_cur * _blockLen > _cto)) { _endOfSamples = true; throw jiterator_error("end of samples!"); }
      std::fill(S.buf.begin(), S.buf.end(), 0.f);
      const long frames = (long)(S.pcm.size() / S.chN), left = frames - S.pos;
      const unsigned readN = left <= 0 ? 0u : (left < (long)S.sampleN ? (unsigned)left : S.sampleN);
      if (readN) memcpy(&S.buf[0], &S.pcm[(size_t)S.pos * S.chN], sizeof(float) * (size_t)readN * S.chN);
      S.pos += readN;
      S.ttl += readN;
      if (readN < S.sampleN) _last = true;
    }
    const size_t offset = (size_t)currentFrame * S.chN * _blockLen;
    for (unsigned i = 0; i < _blockLen; i++) _vector->data[i] = S.buf[offset + (size_t)i * S.chN + _chanX];
    _cur++;
    _increment();
    return _vector;
  }
  // every sample this channel serves from reset() to the end of the stream under lock-step pulling; leaves the shared
  // read position alone
  void whole_stream(std::vector<float>& x) const {
    const Shared& S = shared();
    x.clear();
    if (S.buf.empty()) throw jio_error("no file has been read");
    const long frames = (long)(S.pcm.size() / S.chN);
    long pos = S.cfrom, cur = 0;
    bool last = false;
    while (!(last || (_cto > 0 && cur * (long)_blockLen > _cto))) {
      const long left = frames - pos;
      const long n = left <= 0 ? 0 : (left < (long)S.sampleN ? left : (long)S.sampleN);
      const size_t base = x.size();
      x.resize(base + S.sampleN, 0.f);
      for (long i = 0; i < n; i++) x[base + i] = S.pcm[(size_t)(pos + i) * S.chN + _chanX];
      pos += n;
      cur += S.blockN;
      last = n < (long)S.sampleN;
    }
  }
  bool unread() const { return _frameX < 0; }

 private:
  struct Shared {
    std::vector<float> pcm, buf;
    int chN, samplerate;
    long pos, cfrom;
    unsigned interval, blockN, sampleN, ttl;
    Shared() : chN(0), samplerate(0), pos(0), cfrom(0), interval(30), blockN(0), sampleN(0), ttl(0) {}
  };
  static Shared& shared() { static Shared s; return s; }
  const unsigned _blockLen, _chanX;
  unsigned _firstChanX, _cur;
  bool _last;
  int _cto;
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
    MemorySampleFeature* ms = dynamic_cast<MemorySampleFeature*>(pget(_samp));
    if (ms && ms->shiftLen() == _D && ms->padZeros() && ms->frameX() < 0) { x = ms->samples(); return; }
    // the channels of one interleaved file share their read buffer and advance in lock step: take the whole stream at once
    IterativeSampleFeature* it = dynamic_cast<IterativeSampleFeature*>(pget(_samp));
    if (it && it->size() == _D && it->unread()) { it->whole_stream(x); return; }
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
typedef node_ptr<OverSampledDFTAnalysisBank>::type OverSampledDFTAnalysisBankPtr;

// ---- SubbandBeamformer / SubbandDS / SubbandMVDR ----------------------------------------------------------------
#if !defined(BTKB200_WITH_BTK)
// SnapShotArray (beamformer/spectralinfoarray.h:6-31, beamformer.cc:35-113): the current frame of every channel
// ([C][M], newSample) transposed into per-bin snapshots ([M][C], update / getSnapShot).
class SnapShotArray {
 public:
  SnapShotArray(unsigned fftLn, unsigned nChn) : _fftLen(fftLn), _nChan(nChn) {
    _specSamples = new btk_vector_complex*[_nChan];
    for (unsigned i = 0; i < _nChan; i++) _specSamples[i] = new_cvec(_fftLen);
    _specSnapShots = new btk_vector_complex*[_fftLen];
    for (unsigned i = 0; i < _fftLen; i++) _specSnapShots[i] = new_cvec(_nChan);
  }
  virtual ~SnapShotArray() {
    for (unsigned i = 0; i < _nChan; i++) free_cvec(_specSamples[i]);
    delete[] _specSamples;
    for (unsigned i = 0; i < _fftLen; i++) free_cvec(_specSnapShots[i]);
    delete[] _specSnapShots;
  }
  const btk_vector_complex* getSnapShot(unsigned fbinX) const { return _specSnapShots[fbinX]; }
  void newSample(const btk_vector_complex* samp, unsigned chanX) const {                        // beamformer.cc:76-80
    for (unsigned s = 0; s < _fftLen; s++) {
      _specSamples[chanX]->data[2 * s] = samp->data[2 * s * samp->stride];
      _specSamples[chanX]->data[2 * s + 1] = samp->data[2 * s * samp->stride + 1];
    }
  }
  // beamformer.cc:99-113, its mirror index (fftLen2 - fbinX, not fftLen - fbinX) included
  void newSnapShot(const btk_vector_complex* snapshots, unsigned fbinX) {
    const unsigned fftLen2 = _fftLen / 2;
    memcpy(_specSnapShots[fbinX]->data, snapshots->data, sizeof(double) * 2 * _nChan);
    if (fbinX == 0 || fbinX == fftLen2) return;
    for (unsigned c = 0; c < _nChan; c++) {
      _specSnapShots[fftLen2 - fbinX]->data[2 * c] = snapshots->data[2 * c * snapshots->stride];
      _specSnapShots[fftLen2 - fbinX]->data[2 * c + 1] = -snapshots->data[2 * c * snapshots->stride + 1];
    }
  }
  unsigned fftLen() const { return _fftLen; }
  unsigned nChan() const { return _nChan; }
  virtual void update() {                                                                       // beamformer.cc:82-90
    for (unsigned s = 0; s < _fftLen; s++)
      for (unsigned c = 0; c < _nChan; c++) {
        _specSnapShots[s]->data[2 * c] = _specSamples[c]->data[2 * s];
        _specSnapShots[s]->data[2 * c + 1] = _specSamples[c]->data[2 * s + 1];
      }
  }
  virtual void zero() {
    for (unsigned i = 0; i < _nChan; i++) memset(_specSamples[i]->data, 0, sizeof(double) * 2 * _fftLen);
    for (unsigned i = 0; i < _fftLen; i++) memset(_specSnapShots[i]->data, 0, sizeof(double) * 2 * _nChan);
  }

 protected:
  const unsigned _fftLen, _nChan;
  mutable btk_vector_complex** _specSamples;
  mutable btk_vector_complex** _specSnapShots;
};
typedef std::shared_ptr<SnapShotArray> SnapShotArrayPtr;

// The accessors of beamformerWeights (beamformer/beamformer.h:49-118) that other nodes read: quiescent weights per bin
// (full M bins, upper half = conjugate mirror, beamformer.cc:566-574) and the array manifold.
class beamformerWeights {
 public:
  beamformerWeights(unsigned fftLen, unsigned chanN, bool halfBandShift, unsigned NC = 1)
      : _halfBandShift(halfBandShift), _fftLen(fftLen), _chanN(chanN), _NC(NC) {
    _wq = new btk_vector_complex*[fftLen]; _ta = new btk_vector_complex*[fftLen];
    for (unsigned s = 0; s < fftLen; s++) { _wq[s] = new_cvec(chanN); _ta[s] = new_cvec(chanN); }
  }
  ~beamformerWeights() {
    for (unsigned s = 0; s < _fftLen; s++) { free_cvec(_wq[s]); free_cvec(_ta[s]); }
    delete[] _wq; delete[] _ta;
  }
  unsigned NC() { return _NC; }
  bool isHalfBandShift() { return _halfBandShift; }
  unsigned fftLen() { return _fftLen; }
  unsigned chanN() { return _chanN; }
  btk_vector_complex* wq_f(unsigned fbinX) { return _wq[fbinX]; }
  btk_vector_complex** wq() { return _wq; }
  btk_vector_complex** arrayManifold() { return _ta; }
  // half spectra [B][C] (re, im) -> all M bins
  void fill(const double* wq_half, const double* ta_half) {
    for (unsigned s = 0; s <= _fftLen / 2; s++)
      for (unsigned c = 0; c < _chanN; c++)
        for (int t = 0; t < 2; t++) {
          const double* src = (t ? ta_half : wq_half) + ((size_t)s * _chanN + c) * 2;
          btk_vector_complex** dst = t ? _ta : _wq;
          dst[s]->data[2 * c] = src[0]; dst[s]->data[2 * c + 1] = src[1];
          if (s > 0 && s < _fftLen / 2) { dst[_fftLen - s]->data[2 * c] = src[0]; dst[_fftLen - s]->data[2 * c + 1] = -src[1]; }
        }
  }

 private:
  bool _halfBandShift;
  unsigned _fftLen, _chanN, _NC;
  btk_vector_complex** _wq;
  btk_vector_complex** _ta;
};

// The members of the reference's SubbandBeamformer + SubbandDS that the B200 node builds on
// (beamformer/beamformer.h:126-182, beamformer.cc:1017-1134): channel list, snapshot array, weight objects.
class BeamformerCore : public VectorComplexFeatureStream {
 public:
  BeamformerCore(unsigned fftLen, bool halfBandShift, const String& nm)
      : VectorComplexFeatureStream(fftLen, nm), _fftLen(fftLen), _fftLen2(fftLen / 2), _halfBandShift(halfBandShift) {}
  virtual ~BeamformerCore() { for (size_t i = 0; i < _bfWeightV.size(); i++) delete _bfWeightV[i]; }
  unsigned fftLen() const { return _fftLen; }
  unsigned fftLen2() const { return _fftLen2; }
  unsigned chanN() const { return (unsigned)_channelList.size(); }
  virtual unsigned dim() const { return chanN(); }
  const btk_vector_complex* snapShotArray_f(unsigned fbinX) { return _snapShotArray->getSnapShot(fbinX); }
  virtual SnapShotArrayPtr getSnapShotArray() { return _snapShotArray; }
  void setChannel(VectorComplexFeatureStreamPtr& chan) { _channelList.push_back(chan); }
  virtual void clearChannel() {                                                                 // beamformer.cc:1076-1085
    _channelList.clear();
    for (size_t i = 0; i < _bfWeightV.size(); i++) delete _bfWeightV[i];
    _bfWeightV.clear();
    _snapShotArray.reset();
  }
  virtual void reset() {                                                                        // beamformer.cc:1038-1048
    for (_ChannelIterator it = _channelList.begin(); it != _channelList.end(); ++it) (*it)->reset();
    if (_snapShotArray) _snapShotArray->zero();
    VectorComplexFeatureStream::reset();
  }
  virtual beamformerWeights* getBeamformerWeightObject(unsigned srcX = 0) { return _bfWeightV[srcX]; }

 protected:
  typedef std::list<VectorComplexFeatureStreamPtr> _ChannelList;
  typedef _ChannelList::iterator _ChannelIterator;
  void _allocImage() { if (!_snapShotArray) _snapShotArray.reset(new SnapShotArray(_fftLen, chanN())); }   // :1123-1127
  void _allocBFWeight(int nSource, int NC) {                                                              // :1129-1139
    for (size_t i = 0; i < _bfWeightV.size(); i++) delete _bfWeightV[i];
    _bfWeightV.resize(nSource);
    for (size_t i = 0; i < _bfWeightV.size(); i++) _bfWeightV[i] = new beamformerWeights(_fftLen, chanN(), _halfBandShift, NC);
    _allocImage();
  }
  SnapShotArrayPtr _snapShotArray;
  unsigned _fftLen, _fftLen2;
  _ChannelList _channelList;
  bool _halfBandShift;
  std::vector<beamformerWeights*> _bfWeightV;
};
#else
// inside the reference tree the core IS the reference's SubbandDS: its channel list, its SnapShotArray, its
// beamformerWeights -- reference nodes that take a SubbandDSPtr (ZelinskiPostFilter::setBeamformer, postfilter.h:95-126)
// accept the B200 beamformer and read the snapshots / weights it publishes
typedef ::SubbandDS BeamformerCore;
#endif

// ---- SpectralMatrixArray (beamformer/spectralinfoarray.h:38-54, beamformer.cc:119-163) -----------------------------
// Per bin R <- mu R + (1 - mu) x x^T (NO conjugate, :151-159) on every update().  B200 evaluation: update() records the
// frame, getSpecMatrix() folds everything recorded since the last read on the device in ONE weighted Gram launch
// (btkb200_covariance, tensor cores; frame f of n pending frames gets the weight (1 - mu) mu^(n-1-f)) and adds mu^n times
// the previous matrices.  The device works on the half spectrum: bins above M/2 come back as the element-wise conjugate
// of their mirror bin, which is what Hermitian spectra (an analysis bank's output) give.
class SpectralMatrixArray : public SnapShotArray {
 public:
  SpectralMatrixArray(unsigned fftLn, unsigned nChn, double forgetFact = 0.95)
      : SnapShotArray(fftLn, nChn), _muB200(forgetFact), _R((size_t)(fftLn / 2 + 1) * nChn * nChn * 2, 0.0) {
    _mat = new btk_matrix_complex();
    _mat->size1 = _mat->size2 = _mat->tda = nChn; _mat->block = 0; _mat->owner = 0;
    _mat->data = new double[2 * (size_t)nChn * nChn]();
  }
  virtual ~SpectralMatrixArray() { delete[] _mat->data; delete _mat; }
  virtual void update() {
    SnapShotArray::update();
    const unsigned B = _fftLen / 2 + 1;
    for (unsigned s = 0; s < B; s++)
      for (unsigned c = 0; c < _nChan; c++) {
        _pending.push_back((float)_specSnapShots[s]->data[2 * c * _specSnapShots[s]->stride]);
        _pending.push_back((float)_specSnapShots[s]->data[2 * c * _specSnapShots[s]->stride + 1]);
      }
  }
  virtual void zero() { SnapShotArray::zero(); _pending.clear(); _R.assign(_R.size(), 0.0); }
  // valid until the next getSpecMatrix() / update()
  btk_matrix_complex* getSpecMatrix(unsigned idx) {
    flush();
    const unsigned C = _nChan, h = _fftLen / 2;
    const bool mirror = idx > h;
    const double* src = &_R[(size_t)(mirror ? _fftLen - idx : idx) * C * C * 2];
    for (size_t i = 0; i < (size_t)C * C; i++) { _mat->data[2 * i] = src[2 * i]; _mat->data[2 * i + 1] = mirror ? -src[2 * i + 1] : src[2 * i + 1]; }
    return _mat;
  }

 private:
  void flush() {
    const unsigned B = _fftLen / 2 + 1, C = _nChan;
    const long n = (long)(_pending.size() / ((size_t)2 * B * C));
    if (n == 0) return;
    if (!_plan) _plan.create(_fftLen, 1, 0, 0, C, 0, 0, 1);
    std::vector<double> wt(n), G(_R.size());
    for (long f = 0; f < n; f++) wt[f] = (1.0 - _muB200) * pow(_muB200, (double)(n - 1 - f));
    check(btkb200_covariance(_plan.get(), &_pending[0], n, &wt[0], 0, &G[0]), _plan.get());
    const double keep = pow(_muB200, (double)n);
    for (size_t i = 0; i < _R.size(); i++) _R[i] = keep * _R[i] + G[i];
    _pending.clear();
  }
  double _muB200;
  std::vector<double> _R;          // [B][C][C] (re, im)
  std::vector<float> _pending;     // [n][B][C] complex64
  PlanHandle _plan;
  btk_matrix_complex* _mat;
};

class SubbandBeamformer : public BeamformerCore {
 public:
  SubbandBeamformer(unsigned fftLen = 512, bool halfBandShift = false, const String& nm = "SubbandBeamformer")
      : BeamformerCore(fftLen, halfBandShift, nm), _F(-1) {
    if (halfBandShift) throw j_error("halfBandShift is not supported by the B200 engine");
    _wvec = new_cvec(0);
    _tmp = new_cvec(fftLen);
  }
  ~SubbandBeamformer() { free_cvec(_wvec); free_cvec(_tmp); }
  virtual void clearChannel() { BeamformerCore::clearChannel(); _plan.reset(); _F = -1; }

  virtual const btk_vector_complex* next(int frameX = -5) {
    if (frameX == _frameX && _frameX >= 0) return _vector;
    if (_F < 0) evaluate();
    const int t = _frameX + 1;
    if (t >= _F) { _endOfSamples = true; throw jiterator_error("end of samples!"); }
    const unsigned B = _fftLen / 2 + 1, M = _fftLen;
    const float* half = &_Y[(size_t)t * B * 2];
    for (unsigned s = 0; s < B; s++) { _vector->data[2 * s] = half[2 * s]; _vector->data[2 * s + 1] = half[2 * s + 1]; }
    for (unsigned s = 1; s < M / 2; s++) { _vector->data[2 * (M - s)] = half[2 * s]; _vector->data[2 * (M - s) + 1] = -half[2 * s + 1]; }   // :1189-1194
    publish(t);
    _increment();
    return _vector;
  }
  virtual void reset() { BeamformerCore::reset(); _F = -1; }

  // ---- used by OverSampledDFTSynthesisBank for the fused path
  bool all_own_banks(unsigned& m, unsigned& r, unsigned& dct) {
    if (_channelList.empty()) return false;
    const OverSampledDFTAnalysisBank* a0 = 0;
    for (_ChannelIterator it = _channelList.begin(); it != _channelList.end(); ++it) {
      const OverSampledDFTAnalysisBank* a = dynamic_cast<const OverSampledDFTAnalysisBank*>(pget(*it));
      if (!a || a->frameX() >= 0 || a->_M != _fftLen) return false;
      if (!a0) a0 = a;
      if (a->_m != a0->_m || a->_r != a0->_r || a->_dct != a0->_dct) return false;
    }
    m = a0->_m; r = a0->_r; dct = a0->_dct;
    return true;
  }
  const std::vector<double>& analysis_prototype() {
    return dynamic_cast<const OverSampledDFTAnalysisBank*>(pget(_channelList.front()))->_prototype;
  }
  void interleaved_pcm(std::vector<float>& pcm, long& T) {
    const unsigned C = chanN();
    std::vector<std::vector<float> > xs(C);
    unsigned c = 0;
    T = 0;
    for (_ChannelIterator it = _channelList.begin(); it != _channelList.end(); ++it, ++c) {
      dynamic_cast<OverSampledDFTAnalysisBank*>(pget(*it))->pull_source(xs[c]);
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
  static void fill(btk_vector_complex* v, unsigned n) {
    if (v->size != n) { delete[] v->data; v->data = new double[2 * (size_t)n](); v->size = n; }
  }
  // the snapshots of frame t go into the SnapShotArray other nodes read (SubbandDS::next: newSample per channel, update,
  // beamformer.cc:1152-1161)
  void publish(int t) {
    SnapShotArray* sa = pget(_snapShotArray);
    if (!sa || sa->nChan() != chanN()) return;
    const unsigned C = chanN(), B = _fftLen / 2 + 1, M = _fftLen;
    for (unsigned c = 0; c < C; c++) {
      for (unsigned s = 0; s < B; s++) {
        const float* x = &_snapshots[(((size_t)t * B + s) * C + c) * 2];
        _tmp->data[2 * s] = x[0]; _tmp->data[2 * s + 1] = x[1];
        if (s > 0 && s < M / 2) { _tmp->data[2 * (M - s)] = x[0]; _tmp->data[2 * (M - s) + 1] = -x[1]; }
      }
      sa->newSample(_tmp, c);
    }
    sa->update();
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

  PlanHandle _plan;
  std::vector<float> _snapshots, _Y;
  int _F;
  btk_vector_complex* _wvec;
  btk_vector_complex* _tmp;
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
#if defined(BTKB200_WITH_BTK)
    BeamformerCore::calcArrayManifoldVectors(sampleRate, delays);     // the reference's own weight object, for its nodes
#else
    _allocBFWeight(1, 1);
    sync_weight_object();
#endif
    _F = -1;
  }
  // beamformer.cc:1100-1121 -> calcMainlobe2 / calcMainlobeN (:603-735): unit gain on the target, nulls on the interferers
  virtual void calcArrayManifoldVectors2(double sampleRate, const btk_vector* delaysT, const btk_vector* delaysJ) {
    if (delaysJ->size != chanN())
      throw jdimension_error("The number of delays for an interference signal does not match number of channels (%d vs. %d).\n",
                             (int)delaysJ->size, (int)chanN());
    std::vector<double> dj(delaysJ->size);
    for (size_t i = 0; i < dj.size(); i++) dj[i] = delaysJ->data[i * delaysJ->stride];
    null_weights(sampleRate, delaysT, &dj[0], 2);
#if defined(BTKB200_WITH_BTK)
    BeamformerCore::calcArrayManifoldVectors2(sampleRate, delaysT, delaysJ);
#endif
  }
  virtual void calcArrayManifoldVectorsN(double sampleRate, const btk_vector* delaysT, const btk_matrix* delaysJ, unsigned NC = 2) {
    if (NC < 2 || NC > chanN() || delaysJ->size1 + 1 < NC || delaysJ->size2 != chanN())
      throw jdimension_error("1 < the number of constraints %d <= the number of sensors %d.\n", (int)NC, (int)chanN());
    std::vector<double> dj((size_t)(NC - 1) * chanN());
    for (unsigned n = 0; n + 1 < NC; n++) for (unsigned c = 0; c < chanN(); c++) dj[(size_t)n * chanN() + c] = delaysJ->data[n * delaysJ->tda + c];
    null_weights(sampleRate, delaysT, &dj[0], NC);
#if defined(BTKB200_WITH_BTK)
    BeamformerCore::calcArrayManifoldVectorsN(sampleRate, delaysT, delaysJ, NC);
#endif
  }
  virtual const btk_vector_complex* getWeights(unsigned fbinX) {
    const unsigned C = chanN(), B = _fftLen / 2 + 1;
    std::vector<double> w((size_t)B * C * 2);
    check(btkb200_get_weights(need_plan(), &w[0]), _plan.get());
    fill(_wvec, C);
    memcpy(_wvec->data, &w[(size_t)fbinX * C * 2], sizeof(double) * 2 * C);
    return _wvec;
  }

 protected:
  void null_weights(double sampleRate, const btk_vector* delaysT, const double* dj, unsigned NC) {
    if (delaysT->size != chanN())
      throw jdimension_error("The number of delays does not match number of channels (%d vs. %d).\n", (int)delaysT->size, (int)chanN());
    std::vector<double> d(delaysT->size);
    for (size_t i = 0; i < d.size(); i++) d[i] = delaysT->data[i * delaysT->stride];
    check(btkb200_set_null_weights(need_plan(), sampleRate, &d[0], (unsigned)d.size(), dj, NC), _plan.get());
#if !defined(BTKB200_WITH_BTK)
    _allocBFWeight(1, (int)NC);
    sync_weight_object();
#endif
    _F = -1;
  }
#if !defined(BTKB200_WITH_BTK)
  void sync_weight_object() {
    const unsigned C = chanN(), B = _fftLen / 2 + 1;
    std::vector<double> wq((size_t)B * C * 2), ta((size_t)B * C * 2);
    check(btkb200_get_manifold(_plan.get(), &wq[0]), _plan.get());
    check(btkb200_get_array_manifold(_plan.get(), &ta[0]), _plan.get());
    _bfWeightV[0]->fill(&wq[0], &ta[0]);
  }
#endif
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
typedef node_ptr<SubbandDS>::type SubbandDSPtr;
typedef node_ptr<SubbandMVDR>::type SubbandMVDRPtr;

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
typedef node_ptr<SubbandGSC>::type SubbandGSCPtr;

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
    if (!pget(_bf)) throw j_error("set beamformer's weights \n");   // postfilter.cc:449-452
    if (static_cast<VectorComplexFeatureStream*>(pget(_bf)) != pget(_samp))
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
typedef node_ptr<ZelinskiPostFilter>::type ZelinskiPostFilterPtr;

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
    if (pget(_samp)) _samp->reset();
    VectorFloatFeatureStream::reset();
    _nout = -1;
    _pushed.clear();
  }

 private:
  void need_plan(unsigned C, const double* h) {
    if (!_plan || _plan.C() != C) _plan.create(_M, _m, _r, _dct, C, h, &_prototype[0], _gain);
  }
  void evaluate() {
    SubbandBeamformer* bf = dynamic_cast<SubbandBeamformer*>(pget(_samp));
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
typedef node_ptr<OverSampledDFTSynthesisBank>::type OverSampledDFTSynthesisBankPtr;

}  // namespace btkb200
#endif  // BTKB200_STREAMS_H
