"""Drop-in feature-stream nodes: the reference's Python-visible ``*Ptr`` surface for the hot path, computed by the
sm_100a kernels behind the C ABI (include/btkb200.h).

Mirrors (paths relative to /root/reference/btk):
  FeatureStream protocol  next(frameX=-5) / reset() / size() / isEnd() / current()      stream/stream.h:35-75
  iterator glue           __iter__ = reset() + self, jiterator_error -> StopIteration    modulated/modulated.i:117-173,
                                                                                        include/jexception.i:57-82
  SampleFeature           block source, blockLen/shiftLen/padZeros                       feature/feature.cc:610-659
  OverSampledDFTAnalysisBankPtr(samp, prototype, M, m, r, delayCompensationType)         modulated/modulated.i:117-132
  OverSampledDFTSynthesisBankPtr(samp, prototype, M, m, r, delayCompensationType, gainFactor)   modulated.i:157-173
  SubbandDSPtr(fftLen, halfBandShift) / SubbandMVDRPtr(fftLen, halfBandShift)           beamformer/beamformer.i:171-184, 319-332

The reference pulls one frame per ``next()`` through the whole chain on one CPU thread.  These nodes keep that
pull API but evaluate lazily per UTTERANCE: the first ``next()`` of a node drains its upstream (which may be a plain
Python iterator, like PyVectorComplexFeatureStream allows, btk/stream/pyStream.h:89-130), runs the device kernels
once, and then serves frames from the result.  When a synthesis bank sits on a SubbandDS/SubbandMVDR whose channels
are all analysis banks of this module, the three stages run as ONE fused kernel (btkb200_chain) and no subband data
ever reaches the host.  ``next()`` returns a view of the node's own buffer that the following ``next()`` overwrites
(the reference's non-owning numpy view, include/vector.i:196-211 -- copy it if you keep it).

There is no CPU path here: every number comes from libbtkb200.so; without a GPU the nodes raise BtkError.
"""
from __future__ import annotations

import numpy as np

from . import _capi
from ._capi import BtkError, Plan

FrameResetX = -1


class j_error(Exception):
    """btk/common/jexception.h:41-173 (code JERROR)."""


class jdimension_error(j_error):
    pass


class jconsistency_error(j_error):
    pass


class jiterator_error(StopIteration):
    """End of stream.  The SWIG layer maps it to StopIteration (include/jexception.i:63-65)."""


def _raise(e: BtkError):
    """C-ABI status -> the exception type the reference throws at that point (include/btkb200.h)."""
    if e.code == _capi.EINVAL:
        raise jdimension_error(e.msg) from None
    if e.code == _capi.ESTATE:
        raise j_error(e.msg) from None
    raise e


class FeatureStream:
    """stream/stream.h:35-75."""

    def __init__(self, size: int, name: str):
        self._size, self._name = int(size), name
        self._frameX, self._endOfSamples = FrameResetX, False
        self._vector = None

    def name(self):
        return self._name

    def size(self):
        return self._size

    def frameX(self):
        return self._frameX

    def isEnd(self):
        return self._endOfSamples

    def current(self):
        if self._frameX < 0:
            raise jconsistency_error(f"Frame index ({self._frameX}) < 0.")
        return self.next(self._frameX)

    def reset(self):
        self._frameX, self._endOfSamples = FrameResetX, False

    def next(self, frameX: int = -5):  # pragma: no cover - interface
        raise NotImplementedError

    # Python iterator glue of the SWIG shadow classes (modulated.i:126-128)
    def __iter__(self):
        self.reset()
        return self

    def __next__(self):
        return self.next()


class SampleFeaturePtr(FeatureStream):
    """In-memory block source with SampleFeature::next's rule (feature/feature.cc:610-659): blocks of ``blockLen``
    samples every ``shiftLen``; with ``padZeros`` the last partial block is zero padded, then jiterator_error.
    ``samples`` is one channel (1-D) of un-normalised float samples (feature.cc:273)."""

    def __init__(self, samples=None, blockLen: int = 320, shiftLen: int = 160, padZeros: bool = False,
                 nm: str = "Sample"):
        super().__init__(blockLen, nm)
        self._shift, self._pad = int(shiftLen), bool(padZeros)
        self._samples = np.zeros(0, np.float32) if samples is None else np.ascontiguousarray(samples, np.float32).ravel()
        self._cur = 0
        self._vector = np.zeros(blockLen, np.float32)

    def setSamples(self, samples, sampleRate: float = 16000.0):
        self._samples = np.ascontiguousarray(samples, np.float32).ravel()
        self.reset()

    def samples(self):
        return self._samples

    def reset(self):
        super().reset()
        self._cur = 0

    def next(self, frameX: int = -5):
        if frameX == self._frameX:
            return self._vector
        if frameX >= 0 and frameX - 1 != self._frameX:      # feature.cc:622-625
            raise jconsistency_error(f"Problem in Feature {self._name}: {frameX - 1} != {self._frameX}")
        T, n = self._samples.size, self._size
        if self._cur >= T or (not self._pad and self._cur + n > T):
            self._endOfSamples = True
            raise jiterator_error("end of samples!")
        blk = self._samples[self._cur:self._cur + n]
        self._vector[:blk.size] = blk
        self._vector[blk.size:] = 0.0
        self._cur += self._shift
        self._frameX += 1
        return self._vector


class jio_error(j_error, IOError):
    """common/jexception.h: jio_error (code JIO = 7 -> IOError in Python, include/jexception.i:66-68)."""


class IterativeSampleFeaturePtr(FeatureStream):
    """IterativeSampleFeature (feature/feature.h:301-335, feature.cc:803-896): one node per channel of ONE interleaved
    multichannel file; the nodes share the file image (class-level state, as the reference's static members) and are
    pulled in lock step.  The node of channel ``firstChanX`` refills a 30-s buffer whenever its block counter wraps; the
    stream is therefore a whole number of buffers long, zero padded, and ends at the first wrap after a short read.

    B200 evaluation: read() decodes the file once -- 16-bit PCM goes to the device as raw bytes and is widened there
    (btkb200_convert_pcm, bit-exact; SFC_SET_NORM_FLOAT is off in the reference, feature.cc:849, so samples stay in the
    int16 range).  Analysis banks fed by these nodes take the channel's complete stream in one piece (whole_stream())."""

    _shared = {"pcm": None, "fs": 0, "chN": 0, "pos": 0, "buf": None, "blockN": 0, "sampleN": 0, "ttl": 0, "cfrom": 0}
    _interval = 30

    def __init__(self, chX: int, blockLen: int = 320, firstChanX: int = 0, nm: str = "Iterative Sample"):
        super().__init__(blockLen, nm)
        self._blockLen, self._chanX, self._firstChanX = int(blockLen), int(chX), int(firstChanX)
        self._cur, self._last, self._cto = 0, False, -1
        self._vector = np.zeros(blockLen, np.float32)

    @staticmethod
    def _decode(fileName, samplerate, chN):
        """(raw int16 interleaved, samplerate, channels) of a RIFF/WAVE file with 16-bit PCM, or of a headerless file."""
        import wave
        with open(fileName, "rb") as f:
            head = f.read(12)
        if head[:4] == b"RIFF" and head[8:12] == b"WAVE":
            with wave.open(fileName, "rb") as w:
                if w.getsampwidth() != 2:
                    raise jio_error(f"Could not open file {fileName}: only 16-bit PCM WAVE files are decoded")
                raw = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
                return raw, w.getframerate(), w.getnchannels()
        return np.fromfile(fileName, dtype="<i2"), samplerate, chN

    def read(self, fileName: str, format: int = 0, samplerate: int = 44100, chN: int = 1, cfrom: int = 0, cto: int = -1):
        """feature.cc:826-866."""
        if self._chanX != self._firstChanX:
            return
        try:
            raw, fs, ch = self._decode(fileName, samplerate, chN)
        except OSError as e:
            raise jio_error(f"Could not open file {fileName}.") from e
        self.readRaw(raw, fs, ch, cfrom, cto)

    def readRaw(self, raw_s16, samplerate: int, chN: int, cfrom: int = 0, cto: int = -1):
        """The same with the interleaved 16-bit samples already in memory."""
        if self._chanX != self._firstChanX:
            return
        raw = np.ascontiguousarray(raw_s16, dtype=np.int16).ravel()
        raw = raw[: raw.size // chN * chN]
        try:
            with Plan(64, 1, 0, 1) as plan:
                pcm = plan.convert_pcm(raw, _capi.PCM_S16).reshape(-1, chN)
        except BtkError as e:
            _raise(e)
        if cto > 0 and cto < cfrom:
            raise jconsistency_error(f"Segment cannot start at {cfrom} and end at {cto}")
        S = IterativeSampleFeaturePtr._shared
        S["pcm"], S["fs"], S["chN"], S["pos"], S["cfrom"] = pcm, int(samplerate), int(chN), int(cfrom), int(cfrom)
        S["blockN"] = self._interval * int(samplerate) // self._blockLen + 1
        S["sampleN"] = S["blockN"] * self._blockLen
        S["buf"] = np.zeros((S["sampleN"], chN), np.float32)
        self._cto = cto - cfrom

    def samplesN(self):
        return IterativeSampleFeaturePtr._shared["ttl"]

    def changeFirstChannelID(self, firstChanX: int):
        self._firstChanX = int(firstChanX)

    def reset(self):
        super().reset()
        IterativeSampleFeaturePtr._shared["ttl"] = 0
        self._cur, self._last = 0, False

    def next(self, frameX: int = -5):
        if frameX == self._frameX:
            return self._vector
        S = IterativeSampleFeaturePtr._shared
        if S["pcm"] is None:
            raise jio_error("no file has been read")
        cf = self._cur % S["blockN"]
        if self._chanX == self._firstChanX and cf == 0:
            if self._last or (self._cto > 0 and self._cur * self._blockLen > self._cto):
                raise jiterator_error("end of samples!")
            S["buf"][:] = 0.0
            n = max(0, min(S["sampleN"], S["pcm"].shape[0] - S["pos"]))
            S["buf"][:n] = S["pcm"][S["pos"]:S["pos"] + n]
            S["pos"] += n
            S["ttl"] += n
            if n < S["sampleN"]:
                self._last = True
        self._vector[:] = S["buf"][cf * self._blockLen:(cf + 1) * self._blockLen, self._chanX]
        self._cur += 1
        self._frameX += 1
        return self._vector

    def whole_stream(self) -> np.ndarray:
        """Every sample this channel's node serves from reset() to end of stream when all channels are pulled in lock step
        (the reference's use): the segment from cfrom, zero padded to whole buffers, ending at the first wrap after a short
        read or past cto.  Does not touch the shared read position."""
        S = IterativeSampleFeaturePtr._shared
        if S["pcm"] is None:
            raise jio_error("no file has been read")
        T = S["pcm"].shape[0]
        pos, cur, last, pieces = S["cfrom"], 0, False, []
        while not (last or (self._cto > 0 and cur * self._blockLen > self._cto)):
            n = max(0, min(S["sampleN"], T - pos))
            seg = np.zeros(S["sampleN"], np.float32)
            seg[:n] = S["pcm"][pos:pos + n, self._chanX]
            pieces.append(seg)
            pos += n
            cur += S["blockN"]
            last = n < S["sampleN"]
        return np.concatenate(pieces) if pieces else np.zeros(0, np.float32)


class ChannelExtractionFeaturePtr(FeatureStream):
    """ChannelExtractionFeature (feature/feature.h:1823-1841, feature.cc:3885-3900): channel chX of chN out of a stream of
    interleaved blocks; output block i = input[i * chN + chX]."""

    def __init__(self, src, chX: int = 0, chN: int = 1, nm: str = "ChannelExtraction"):
        if chX >= chN or src.size() % chN != 0:
            raise jdimension_error(f"channel {chX} of {chN} out of blocks of {src.size()}")
        super().__init__(src.size() // chN, nm)
        self._src, self._chX, self._chN = src, int(chX), int(chN)
        self._vector = np.zeros(self._size, np.float32)

    def reset(self):
        self._src.reset()
        super().reset()

    def next(self, frameX: int = -5):
        if frameX == self._frameX:
            return self._vector
        if frameX >= 0 and frameX - 1 != self._frameX:
            raise jconsistency_error(f"Problem in Feature {self._name}: {frameX - 1} != {self._frameX}")
        self._frameX += 1
        allch = np.asarray(self._src.next(self._frameX), np.float32)
        self._vector[:] = allch[self._chX::self._chN][: self._size]
        return self._vector


class Conversion24bit2FloatPtr(FeatureStream):
    """Conversion24bit2Float (feature/feature.h:148-158, feature.cc:190-217): packed big-endian 24-bit samples (the
    Mark-III/IV wire format, 3 bytes per sample) -> float.  B200 evaluation: the byte stream is drained once and widened on
    the device in one call (btkb200_convert_pcm, BTKB200_PCM_S24BE; bit-exact), then served block by block."""

    def __init__(self, src, nm: str = "Conversion from 24 bit integer to Float"):
        super().__init__(src.size() // 3, nm)
        self._src = src
        self._all = None
        self._vector = np.zeros(self._size, np.float32)

    def reset(self):
        self._src.reset()
        super().reset()
        self._all = None

    def next(self, frameX: int = -5):
        if frameX == self._frameX:
            return self._vector
        if frameX >= 0 and frameX - 1 != self._frameX:
            raise jconsistency_error(f"Problem in Feature {self._name}: {frameX - 1} != {self._frameX}")
        if self._all is None:
            blocks = _drain(self._src)
            raw = np.concatenate([np.asarray(b).astype(np.uint8) for b in blocks]) if blocks else np.zeros(0, np.uint8)
            raw = raw[: raw.size // 3 * 3].reshape(-1, 3)
            try:
                with Plan(64, 1, 0, 1) as plan:
                    self._all = plan.convert_pcm(raw, _capi.PCM_S24BE) if raw.size else np.zeros(0, np.float32)
            except BtkError as e:
                _raise(e)
        t = self._frameX + 1
        if (t + 1) * self._size > self._all.size:
            self._endOfSamples = True
            raise jiterator_error("end of samples!")
        self._vector[:] = self._all[t * self._size:(t + 1) * self._size]
        self._frameX = t
        return self._vector


def _drain(stream) -> list:
    """Pull every remaining frame of an upstream node or plain iterator (copies: upstream reuses its buffer)."""
    frames = []
    if isinstance(stream, FeatureStream):
        while True:
            try:
                frames.append(np.array(stream.next(), copy=True))
            except StopIteration:
                break
    else:
        for f in stream:
            frames.append(np.array(f, copy=True))
    return frames


class _FilterBank:
    def _geometry(self, prototype, M, m, r, dct):
        self._M, self._m, self._r, self._dct = int(M), int(m), int(r), int(dct)
        self._R = 1 << self._r
        self._D = self._M // self._R
        proto = np.ascontiguousarray(prototype, np.float64).ravel()
        if proto.size != self._M * self._m:      # modulated.cc:269-271
            raise jconsistency_error(f"Prototype sizes do not match ({proto.size} vs. {self._M * self._m}).")
        self._prototype = proto.copy()           # copied at construction (modulated.cc:275-276)

    def fftLen(self):
        return self._M

    def polyphase(self, m, n):
        return float(self._prototype[m + self._M * n])     # modulated.h:233-236


class OverSampledDFTAnalysisBankPtr(FeatureStream, _FilterBank):
    """modulated/modulated.cc:359-516.  ``next()`` yields complex128[M] full Hermitian spectra."""

    def __init__(self, samp, prototype, M: int = 256, m: int = 3, r: int = 0, delayCompensationType: int = 0,
                 nm: str = "OverSampledDFTAnalysisBank"):
        FeatureStream.__init__(self, M, nm)
        self._geometry(prototype, M, m, r, delayCompensationType)
        self._samp = samp
        self._frames = None
        self._plan = None
        self._vector = np.zeros(M, np.complex128)

    def nBlocks(self):
        return 4

    def subSampRate(self):
        return 2

    # -- hooks for the fused path
    def _source_samples(self):
        """All samples of this bank's source as one float32 array, or None when the source is not a block source
        with blockLen == shiftLen == D (then frames have to be pulled one by one)."""
        s = self._samp
        if isinstance(s, SampleFeaturePtr) and s.size() == self._D and s._shift == self._D and s._pad:
            return s.samples()
        if isinstance(s, IterativeSampleFeaturePtr) and s.size() == self._D and s._frameX == FrameResetX:
            # the channels of one interleaved file share their read buffer and have to advance in lock step; a bank takes
            # its channel's complete stream in one piece instead
            return s.whole_stream()
        return None

    def _pull_source(self) -> np.ndarray:
        direct = self._source_samples()
        if direct is not None:
            return direct
        blocks = _drain(self._samp)          # any exception from the source ends the input (modulated.cc:493-501)
        for b in blocks:
            if b.size != self._D:
                raise jdimension_error(f"Input block length ({b.size}) != D ({self._D})")
        return np.concatenate(blocks).astype(np.float32) if blocks else np.zeros(0, np.float32)

    def _evaluate(self):
        x = self._pull_source()
        try:
            if self._plan is None:
                self._plan = Plan(self._M, self._m, self._r, 1, h=self._prototype, dct=self._dct)
            snap = self._plan.analysis(x[:, None])[:, :, 0]           # [F][B]
        except BtkError as e:
            _raise(e)
        self._frames = snap

    def next(self, frameX: int = -5):
        if frameX == self._frameX and self._frameX >= 0:
            return self._vector
        if self._frames is None:
            self._evaluate()
        t = self._frameX + 1
        if t >= self._frames.shape[0]:
            self._endOfSamples = True
            raise jiterator_error("end of samples!")
        half = self._frames[t]
        B = self._M // 2 + 1
        self._vector[:B] = half
        self._vector[B:] = np.conj(half[1:self._M // 2][::-1])       # X[M-s] = conj X[s]
        self._frameX = t
        return self._vector

    def reset(self):
        FeatureStream.reset(self)
        if isinstance(self._samp, FeatureStream):
            self._samp.reset()
        self._frames = None


class SnapShotArrayPtr:
    """SnapShotArray (beamformer/spectralinfoarray.h:6-31, beamformer.cc:35-113; SWIG shadow beamformer.i:76-103): the
    current frame of every channel ([C][M], newSample) transposed into per-bin snapshots ([M][C], update/getSnapShot)."""

    def __init__(self, fftLn: int, nChn: int):
        self._fftLen, self._nChan = int(fftLn), int(nChn)
        self._samples = np.zeros((self._nChan, self._fftLen), np.complex128)
        self._snapshots = np.zeros((self._fftLen, self._nChan), np.complex128)

    def fftLen(self):
        return self._fftLen

    def nChan(self):
        return self._nChan

    def newSample(self, samp, chanX: int):
        v = np.asarray(samp, np.complex128).ravel()
        if v.size != self._fftLen or not 0 <= chanX < self._nChan:
            raise jdimension_error(f"newSample: {v.size} bins for channel {chanX} of a {self._fftLen} x {self._nChan} array")
        self._samples[chanX] = v

    def newSnapShot(self, snapshots, fbinX: int):
        """beamformer.cc:99-113, its mirror index (fftLen2 - fbinX) included."""
        v = np.asarray(snapshots, np.complex128).ravel()
        h = self._fftLen // 2
        self._snapshots[fbinX] = v
        if fbinX != 0 and fbinX != h:
            self._snapshots[h - fbinX] = np.conj(v)

    def getSnapShot(self, fbinX: int):
        return self._snapshots[fbinX]

    def update(self):
        self._snapshots[:] = self._samples.T

    def zero(self):
        self._samples[:] = 0
        self._snapshots[:] = 0


class SpectralMatrixArrayPtr(SnapShotArrayPtr):
    """SpectralMatrixArray (spectralinfoarray.h:38-54, beamformer.cc:119-163; SWIG shadow beamformer.i:105-114): per bin
    R <- mu R + (1 - mu) x x^T (no conjugate, :151-159) on every update().

    B200 evaluation: update() only records the frame; getSpecMatrix() folds everything recorded since the last read on
    the device in ONE weighted Gram launch (btkb200_covariance, tensor cores; frame f of n pending gets the weight
    (1 - mu) mu^(n-1-f)) and adds mu^n times the previous matrices.  The device works on the half spectrum: bins above
    M/2 are returned as the element-wise conjugate of their mirror, which is what the analysis bank's Hermitian spectra
    give (for spectra that are not Hermitian the upper bins differ from the reference's)."""

    def __init__(self, fftLn: int, nChn: int, forgetFact: float = 0.95):
        super().__init__(fftLn, nChn)
        self._mu = float(forgetFact)
        self._B = self._fftLen // 2 + 1
        self._R = np.zeros((self._B, self._nChan, self._nChan), np.complex128)
        self._pending = []
        self._plan = None

    def update(self):
        super().update()
        self._pending.append(self._snapshots[:self._B].astype(np.complex64))

    def zero(self):
        super().zero()
        self._R[:] = 0
        self._pending = []

    def _flush(self):
        n = len(self._pending)
        if n == 0:
            return
        if self._plan is None:
            self._plan = Plan(self._fftLen, 1, 0, self._nChan)
        wt = (1.0 - self._mu) * self._mu ** np.arange(n - 1, -1, -1, dtype=np.float64)
        try:
            G = self._plan.covariance(np.ascontiguousarray(np.stack(self._pending)), wt, conjugate=False)
        except BtkError as e:
            _raise(e)
        self._R = self._mu ** n * self._R + G
        self._pending = []

    def getSpecMatrix(self, idx: int):
        self._flush()
        h = self._fftLen // 2
        return self._R[idx] if idx <= h else np.conj(self._R[self._fftLen - idx])


class _SubbandBeamformer(FeatureStream):
    """SubbandBeamformer: ordered channel list, snapshot array, weight apply (beamformer.cc:1017-1048, 1137-1200)."""

    def __init__(self, fftLen: int, halfBandShift: bool, nm: str):
        super().__init__(fftLen, nm)
        if halfBandShift:
            raise j_error("halfBandShift is not supported by the B200 engine")
        self._fftLen = int(fftLen)
        self._channels = []
        self._plan = None
        self._Y = None
        self._snap = None
        self._vector = np.zeros(fftLen, np.complex128)

    def setChannel(self, chan):
        self._channels.append(chan)

    def clearChannel(self):
        self._channels = []
        self._drop_plan()

    def chanN(self):
        return len(self._channels)

    def fftLen(self):
        return self._fftLen

    def dim(self):
        return self._fftLen

    def _drop_plan(self):
        if self._plan is not None:
            self._plan.close()
        self._plan = None

    def _need_plan(self) -> Plan:
        C = self.chanN()
        if C == 0:
            raise j_error("No channel is set")
        if self._plan is None or self._plan.C != C:
            self._drop_plan()
            # geometry of the analysis side when the channels are our banks (needed for the staged analysis call)
            a = self._channels[0]
            if isinstance(a, OverSampledDFTAnalysisBankPtr):
                self._plan = Plan(self._fftLen, a._m, a._r, C, h=a._prototype, dct=a._dct)
            else:
                self._plan = Plan(self._fftLen, 1, 0, C)
        return self._plan

    def calcArrayManifoldVectors(self, sampleRate: float, delays):
        """beamformer.cc:1087-1091 -> beamformerWeights::calcMainlobe (:531-594)."""
        d = np.ascontiguousarray(delays, np.float64).ravel()
        if d.size != self.chanN():
            raise jdimension_error(f"Number of delays does not match number of channels ({d.size} vs. {self.chanN()}).")
        try:
            self._need_plan().set_ds_weights(sampleRate, d)
        except BtkError as e:
            _raise(e)
        self._weights_changed()

    def calcArrayManifoldVectors2(self, sampleRate: float, delaysT, delaysJ):
        """beamformer.cc:1100-1104 -> calcMainlobe2 (:603-623): unit gain on the target, a null on one interferer."""
        self.calcArrayManifoldVectorsN(sampleRate, delaysT, np.atleast_2d(np.asarray(delaysJ, np.float64)), 2)

    def calcArrayManifoldVectorsN(self, sampleRate: float, delaysT, delaysJ, NC: int = 2):
        """beamformer.cc:1113-1121 -> calcMainlobeN (:632-735); delaysJ is [NC-1][C]."""
        dT = np.ascontiguousarray(delaysT, np.float64).ravel()
        dJ = np.atleast_2d(np.asarray(delaysJ, np.float64))
        C = self.chanN()
        if NC < 2 or NC > C or dJ.shape[0] < NC - 1:
            raise jdimension_error(f"1 < the number of constraints {NC} <= the number of sensors {C}.")
        if dT.size != C or dJ.shape[1] != C:
            raise jdimension_error(f"The number of delays does not match number of channels ({dT.size} vs. {C}).")
        try:
            self._need_plan().set_null_weights(sampleRate, dT, dJ[:NC - 1])
        except BtkError as e:
            _raise(e)
        self._weights_changed()

    def getSnapShotArray(self):
        """SubbandBeamformer::getSnapShotArray (beamformer.h:141): the snapshots of the current frame, all M bins."""
        if self._snap is None or self._frameX < 0:
            raise j_error("no snapshot yet")
        sa = SnapShotArrayPtr(self._fftLen, self.chanN())
        half = self._snap[self._frameX].astype(np.complex128)            # [B][C]
        M = self._fftLen
        sa._snapshots[:M // 2 + 1] = half
        sa._snapshots[M // 2 + 1:] = np.conj(half[1:M // 2][::-1])
        sa._samples[:] = sa._snapshots.T
        return sa

    def _weights_changed(self):
        self._Y = None

    def getWeights(self, fbinX: int):
        try:
            return self._need_plan().get_weights()[fbinX]
        except BtkError as e:
            _raise(e)

    def _all_own_banks(self) -> bool:
        chans = self._channels
        return bool(chans) and all(isinstance(c, OverSampledDFTAnalysisBankPtr) and c._frameX == FrameResetX and
                                   (c._M, c._m, c._r, c._dct) == (chans[0]._M, chans[0]._m, chans[0]._r, chans[0]._dct)
                                   and c._M == self._fftLen for c in chans)

    def _interleaved_pcm(self):
        """[T][C] float32 from the channels' sources when every channel is one of our analysis banks."""
        xs = [c._pull_source() for c in self._channels]
        T = max(x.size for x in xs)
        pcm = np.zeros((T, len(xs)), np.float32)
        for c, x in enumerate(xs):
            pcm[:x.size, c] = x
        return pcm

    def _check_weights(self):
        if not self._need_plan().has_weights():
            raise j_error("call calcArrayManifoldVectorsX() once")   # beamformer.cc:1140-1143

    def _evaluate(self):
        self._check_weights()
        p = self._plan
        try:
            if self._all_own_banks():
                snap = p.analysis(self._interleaved_pcm())                     # one multichannel launch
            else:
                per = [_drain(c) for c in self._channels]
                F = min(len(f) for f in per)
                B = self._fftLen // 2 + 1
                snap = np.empty((F, B, len(per)), np.complex64)
                for c, fr in enumerate(per):
                    snap[:, :, c] = np.asarray(fr[:F])[:, :B]
            self._snap = snap
            self._Y = p.beamform(snap)
        except BtkError as e:
            _raise(e)

    def snapShotArray_f(self, fbinX: int):
        if self._snap is None or self._frameX < 0:
            raise j_error("no snapshot yet")
        return self._snap[self._frameX, fbinX].astype(np.complex128)

    def next(self, frameX: int = -5):
        if frameX == self._frameX and self._frameX >= 0:
            return self._vector
        if self._Y is None:
            self._evaluate()
        t = self._frameX + 1
        if t >= self._Y.shape[0]:
            self._endOfSamples = True
            raise jiterator_error("end of samples!")
        half = self._Y[t]
        B, M = self._fftLen // 2 + 1, self._fftLen
        self._vector[:B] = half
        self._vector[B:] = np.conj(half[1:M // 2][::-1])       # beamformer.cc:1189-1194
        self._frameX = t
        return self._vector

    def reset(self):
        super().reset()
        for c in self._channels:
            if isinstance(c, FeatureStream):
                c.reset()
        self._Y = None
        self._snap = None


class SubbandDSPtr(_SubbandBeamformer):
    """beamformer/beamformer.h:159-182."""

    def __init__(self, fftLen: int = 512, halfBandShift: bool = False, nm: str = "SubbandDS"):
        super().__init__(fftLen, halfBandShift, nm)


class SubbandGSCPtr(_SubbandBeamformer):
    """beamformer/beamformer.h:186-210, beamformer.cc:1296-1447 with FIXED active weights (the adaptive subclasses are out of
    scope): calcGSCWeights -> setActiveWeights_f per bin -> next().  The effective weights wq - B wa are installed when
    the utterance is evaluated."""

    def __init__(self, fftLen: int = 512, halfBandShift: bool = False, nm: str = "SubbandGSC"):
        super().__init__(fftLen, halfBandShift, nm)
        self._normalize = False
        self._gsc_dirty = False

    def normalizeWeight(self, flag: bool):
        self._normalize = bool(flag)
        self._gsc_dirty = True
        self._weights_changed()

    def calcGSCWeights(self, sampleRate: float, delaysT):
        d = np.ascontiguousarray(delaysT, np.float64).ravel()
        if d.size != self.chanN():
            raise jdimension_error(f"Number of delays does not match number of channels ({d.size} vs. {self.chanN()}).")
        try:
            self._need_plan().gsc_calc_weights(sampleRate, d)
        except BtkError as e:
            _raise(e)
        self._gsc_dirty = True
        self._weights_changed()

    def setActiveWeights_f(self, fbinX: int, packedWeight):
        try:
            if fbinX <= self._fftLen // 2:          # the upper bins are the conjugate mirror and never read (:1326-1352)
                self._need_plan().gsc_set_active_weights(fbinX, packedWeight)
        except BtkError as e:
            _raise(e)
        self._gsc_dirty = True
        self._weights_changed()

    def zeroActiveWeights(self):
        try:
            self._need_plan().gsc_zero_active_weights()
        except BtkError as e:
            _raise(e)
        self._gsc_dirty = True
        self._weights_changed()

    def getBlockingMatrix(self, srcX: int, fbinX: int):
        try:
            return self._need_plan().gsc_blocking_matrix(fbinX)
        except BtkError as e:
            _raise(e)

    def _check_weights(self):
        if self._gsc_dirty:
            try:
                self._need_plan().gsc_apply(self._normalize)
            except BtkError as e:
                if e.code == _capi.ESTATE:
                    raise j_error("call calcGSCWeightsX() once") from None      # beamformer.cc:1307-1310
                _raise(e)
            self._gsc_dirty = False
        if not self._need_plan().has_weights():
            raise j_error("call calcGSCWeightsX() once")


class SubbandMVDRPtr(_SubbandBeamformer):
    """beamformer/beamformer.h:333-388, beamformer.cc:2321-2635."""

    def __init__(self, fftLen: int = 512, halfBandShift: bool = False, nm: str = "SubbandMVDR"):
        super().__init__(fftLen, halfBandShift, nm)

    def setNoiseSpatialSpectralMatrix(self, fbinX: int, Rnn) -> bool:
        R = np.asarray(Rnn)
        if R.shape != (self.chanN(), self.chanN()):     # the reference prints and returns false (:2457-2464)
            return False
        self._need_plan().set_covariance(fbinX, R)
        return True

    def getNoiseSpatialSpectralMatrix(self, fbinX: int):
        try:
            return self._need_plan().get_covariance(fbinX)
        except BtkError as e:
            _raise(e)

    def setDiffuseNoiseModel(self, micPositions, sampleRate: float, sspeed: float = 343740.0) -> bool:
        mp = np.asarray(micPositions, np.float64)
        if mp.ndim != 2 or mp.shape[0] != self.chanN() or mp.shape[1] < 3:
            return False
        self._need_plan().set_diffuse_noise_model(mp, sampleRate, sspeed)
        return True

    def setAllLevelsOfDiagonalLoading(self, diagonalWeight: float):
        try:
            self._need_plan().diag_load(float(diagonalWeight))
        except BtkError as e:
            _raise(e)

    def setLevelOfDiagonalLoading(self, fbinX: int, diagonalWeight: float):
        try:
            self._need_plan().diag_load(float(diagonalWeight), fbinX)
        except BtkError as e:
            _raise(e)

    def divideAllNonDiagonalElements(self, myu: float):
        try:
            self._need_plan().divide_nondiagonal(float(myu))
        except BtkError as e:
            _raise(e)

    def calcMVDRWeights(self, sampleRate: float, dThreshold: float = 1.0e-8, calcInverseMatrix: bool = True) -> bool:
        try:
            self._need_plan().solve_mvdr(sampleRate, dThreshold)
        except BtkError as e:
            _raise(e)
        self._mvdr_ready = True
        self._weights_changed()
        return True

    def _check_weights(self):
        # SubbandMVDR::next (beamformer.cc:2587-2594): the manifold first, then the MVDR weights -- delay-and-sum output is
        # never served in place of MVDR output
        if not self._need_plan().has_weights():
            raise j_error("call calcArrayManifoldVectorsX() once")
        if not getattr(self, "_mvdr_ready", False):
            raise j_error("call calcMVDRWeights() once")

    def getMVDRWeights(self, fbinX: int):
        return self.getWeights(fbinX)


class AnalysisOversampledDFTDesignPtr:
    """modulated/modulated.i shadow of AnalysisOversampledDFTDesign (prototypeDesign.h:148-166): design() / calcError() / save()."""

    def __init__(self, M: int = 512, m: int = 2, r: int = 1, wpFactor: float = 1.0, tau_h: int = -1):
        self._M, self._m, self._r, self._wp, self._tau = int(M), int(m), int(r), float(wpFactor), int(tau_h)
        self._proto = None
        self._err = None

    def design(self, tolerance: float = 1.0e-07):
        try:
            self._proto, self._err = _capi.design_analysis_prototype(self._M, self._m, self._r, self._wp, self._tau, tolerance)
        except BtkError as e:
            _raise(e)
        return self._proto

    def calcError(self, doPrint: bool = True):
        if self._err is None:
            raise j_error("call design() first")
        if doPrint:
            print("eps_p = %f\neps_i = %f" % (self._err[0], self._err[1]))
        return self._err

    def save(self, fileName: str):                    # one "%24.21e" per line (prototypeDesign.cc:214-220)
        with open(fileName, "w") as fp:
            for x in self._proto:
                fp.write("%24.21e\n" % x)


class SynthesisOversampledDFTDesignPtr(AnalysisOversampledDFTDesignPtr):
    """prototypeDesign.h:177-218: SynthesisOversampledDFTDesign(h, M, m, r, v, wpFactor, tau_g)."""

    def __init__(self, h, M: int = 512, m: int = 2, r: int = 1, v: float = 1.0, wpFactor: float = 1.0, tau_g: int = -1):
        super().__init__(M, m, r, wpFactor, tau_g)
        self._h, self._v = np.ascontiguousarray(h, np.float64).ravel().copy(), float(v)

    def design(self, tolerance: float = 1.0e-07):
        try:
            self._proto, self._err = _capi.design_synthesis_prototype(self._h, self._M, self._m, self._r, self._v, self._wp,
                                                                      self._tau, tolerance)
        except BtkError as e:
            _raise(e)
        return self._proto

    def calcError(self, doPrint: bool = True):
        if self._err is None:
            raise j_error("call design() first")
        if doPrint:
            print("eps_t = %f\neps_r = %f" % (self._err[0], self._err[1]))
        return self._err


class AnalysisNyquistMDesignPtr(AnalysisOversampledDFTDesignPtr):
    """modulated/modulated.i:366-385 shadow of AnalysisNyquistMDesign (prototypeDesign.h:221-235, .cc:955-1001)."""

    def design(self, tolerance: float = 1.0e-07):
        try:
            self._proto, self._path = _capi.design_analysis_nyquist(self._M, self._m, self._r, self._wp, self._tau, tolerance)
        except BtkError as e:
            _raise(e)
        return self._proto

    def solutionPath(self) -> int:
        """3 or 4: which of the reference's "alternate solutions" the last design() took (it prints the number)."""
        return self._path

    def calcError(self, doPrint: bool = True):
        raise j_error("calcError() is not provided for the Nyquist(M) designs")


class SynthesisNyquistMDesignPtr(AnalysisNyquistMDesignPtr):
    """prototypeDesign.h:242-264, .cc:1003-1119: SynthesisNyquistMDesign(h, M, m, r, wpFactor, tau_g)."""

    def __init__(self, h, M: int = 512, m: int = 2, r: int = 1, wpFactor: float = 1.0, tau_g: int = -1):
        super().__init__(M, m, r, wpFactor, tau_g)
        self._h = np.ascontiguousarray(h, np.float64).ravel().copy()

    def design(self, tolerance: float = 1.0e-07):
        try:
            self._proto, self._path = _capi.design_synthesis_nyquist(self._h, self._M, self._m, self._r, self._wp, self._tau, tolerance)
        except BtkError as e:
            _raise(e)
        return self._proto


TYPE_ZELINSKI1_REAL, TYPE_ZELINSKI1_ABS, TYPE_APAB, TYPE_ZELINSKI2, NO_USE_POST_FILTER = 0x01, 0x02, 0x04, 0x08, 0x00


class ZelinskiPostFilterPtr(FeatureStream):
    """postfilter/postfilter.h:95-126, postfilter.cc:340-500: the node both shipped drivers put between the beamformer and
    the synthesis bank (src/beamformerDS.cc:154,183).  ``output`` must be the SubbandDS/SubbandMVDR node that is also
    given to setBeamformer(); its snapshots and array manifold feed the filter (setSnapShotArray / setArrayManifoldVector
    with foreign sources are not supported by the B200 engine)."""

    def __init__(self, output, fftLen: int, alpha: float = 0.6, type: int = 2, minFrames: int = 0,
                 nm: str = "ZelinskPostFilter"):
        super().__init__(fftLen, nm)
        if output.size() != fftLen:          # postfilter.cc:355-358
            raise jdimension_error(f"Input block length ({output.size()}) != fftLen ({fftLen})")
        self._samp, self._alpha, self._type, self._minFrames = output, float(alpha), int(type), int(minFrames)
        self._bf = None
        self._Y = None
        self._W = None
        self._vector = np.zeros(fftLen, np.complex128)

    def setBeamformer(self, beamformer):
        self._bf = beamformer

    def getPostFilterWeights(self):
        """wp1 of the frame last returned (full M bins, conjugate-mirrored like postfilter.cc:185-186)."""
        if self._W is None or self._frameX < 0:
            return None
        M = self._size
        half = self._W[self._frameX].astype(np.complex128)
        return np.concatenate([half, np.conj(half[1:M // 2][::-1])])

    def _evaluate(self):
        bf = self._bf
        if bf is None or not isinstance(bf, _SubbandBeamformer):
            raise j_error("set beamformer's weights \n")            # postfilter.cc:449-452
        if bf is not self._samp:
            raise j_error("the B200 post-filter expects its input stream to be the beamformer given to setBeamformer()")
        if bf._snap is None:
            bf._evaluate()
        try:
            self._Y, self._W = bf._plan.beamform_zelinski(bf._snap, self._alpha, self._type, self._minFrames)
        except BtkError as e:
            _raise(e)

    def next(self, frameX: int = -5):
        if frameX == self._frameX and self._frameX >= 0:
            return self._vector
        if self._Y is None:
            self._evaluate()
        t = self._frameX + 1
        if t >= self._Y.shape[0]:
            self._endOfSamples = True
            raise jiterator_error("end of samples!")
        half = self._Y[t]
        M = self._size
        B = M // 2 + 1
        self._vector[:B] = half
        self._vector[B:] = np.conj(half[1:M // 2][::-1])       # postfilter.cc:213-216
        self._frameX = t
        return self._vector

    def reset(self):
        super().reset()
        if isinstance(self._samp, FeatureStream):
            self._samp.reset()
        self._Y = None
        self._W = None


class OverSampledDFTSynthesisBankPtr(FeatureStream, _FilterBank):
    """modulated/modulated.cc:521-674.  ``next()`` yields float32[D] blocks (scaled 1/D like the reference)."""

    def __init__(self, samp, prototype, M: int = 256, m: int = 3, r: int = 0, delayCompensationType: int = 0,
                 gainFactor: int = 1, nm: str = "OverSampledDFTSynthesisBank"):
        D = int(M) >> int(r)
        FeatureStream.__init__(self, D, nm)
        self._geometry(prototype, M, m, r, delayCompensationType)
        self._samp, self._gain = samp, int(gainFactor)
        self._out = None
        self._plan = None
        self._fused = False
        self._vector = np.zeros(D, np.float32)

    def fused(self) -> bool:
        """True when the last evaluation ran analysis -> beamformer -> synthesis as one kernel."""
        return self._fused

    def _can_fuse(self) -> bool:
        bf = self._samp
        return (isinstance(bf, _SubbandBeamformer) and bf._frameX == FrameResetX and bf._all_own_banks()
                and bf._fftLen == self._M and bf._channels[0]._m == self._m and bf._channels[0]._r == self._r
                and bf._channels[0]._dct == self._dct)

    def _evaluate(self):
        try:
            if self._can_fuse():
                bf = self._samp
                bf._check_weights()
                C = bf.chanN()
                if self._plan is None or self._plan.C != C:
                    if self._plan is not None:
                        self._plan.close()
                    self._plan = Plan(self._M, self._m, self._r, C, h=bf._channels[0]._prototype, g=self._prototype,
                                      dct=self._dct, gain=self._gain)
                self._plan.set_weights(bf._plan.get_weights())
                out = self._plan.chain(bf._interleaved_pcm())
                self._fused = True
            else:
                frames = _drain(self._samp)
                B = self._M // 2 + 1
                Y = np.asarray(frames, np.complex64)[:, :B] if frames else np.zeros((0, B), np.complex64)
                if self._plan is None or self._plan.C != 1:
                    self._plan = Plan(self._M, self._m, self._r, 1, g=self._prototype, dct=self._dct, gain=self._gain)
                out = self._plan.synthesis(Y)
                self._fused = False
        except BtkError as e:
            _raise(e)
        self._out = out.reshape(-1, self._D)

    def next(self, frameX: int = -5):
        if frameX == self._frameX and self._frameX >= 0:
            return self._vector
        if self._out is None:
            self._evaluate()
        t = self._frameX + 1
        if t >= self._out.shape[0]:
            self._endOfSamples = True
            raise jiterator_error("end of samples!")
        self._vector[:] = self._out[t]
        self._frameX = t
        return self._vector

    def reset(self):
        FeatureStream.reset(self)
        if isinstance(self._samp, FeatureStream):
            self._samp.reset()
        self._out = None
