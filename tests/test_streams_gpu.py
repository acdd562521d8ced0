"""GPU tier: the drop-in stream nodes (distantspeechrecognition-mirror_b200/streams.py) used the way the reference's
own scripts and drivers use the *Ptr classes, checked against the oracle.

  round trip     btk/tools/filterbank/testNyquistFilterBankDesign.py:45-66
  MVDR chain     btk/src/superdirectiveBeamformer.cc:150-220 (without the Zelinski post-filter)
"""
import copy

import numpy as np
import pytest

import btk_b200
import btk_oracle as bo
from btk_b200 import (OverSampledDFTAnalysisBankPtr, OverSampledDFTSynthesisBankPtr, SampleFeaturePtr, SubbandDSPtr,
                      SubbandMVDRPtr)
from btk_b200.streams import j_error, jconsistency_error, jdimension_error
from conftest import proto

pytestmark = pytest.mark.gpu
wl = btk_b200.workloads
FS = 16000.0


def test_filterbank_round_trip_like_reference_script(prototypes):
    M, m, r = 256, 4, 1
    D = M >> r
    h, g = proto(prototypes, M, m, r)
    x = wl.noise_recording(20000, 1, seed=3)[:, 0]
    sampleFB = SampleFeaturePtr(x, blockLen=D, shiftLen=D, padZeros=True)
    analysisFB = OverSampledDFTAnalysisBankPtr(sampleFB, prototype=h, M=M, m=m, r=r, delayCompensationType=2)
    synthesisFB = OverSampledDFTSynthesisBankPtr(analysisFB, prototype=g, M=M, m=m, r=r, delayCompensationType=2)
    wavebuffer = []
    for b in synthesisFB:
        wavebuffer.append(copy.deepcopy(b))          # the script copies: next() reuses its buffer
    out = np.concatenate(wavebuffer)
    geo = bo.BankGeometry(M, m, r, 2)
    X = bo.analysis(x, h, geo)
    ref = bo.synthesis(X, g, geo).reshape(-1)
    assert out.shape == ref.shape
    assert bo.snr_db(out, ref) >= 70.0
    assert synthesisFB.isEnd()
    # reset() rewinds the whole chain (modulated.cc:454-459, 666-674)
    again = np.concatenate([copy.deepcopy(b) for b in synthesisFB])
    assert np.array_equal(again, out)


def test_analysis_frames_and_protocol(prototypes):
    M, m, r = 512, 2, 2
    h, _ = proto(prototypes, M, m, r)
    x = wl.noise_recording(5000, 1, seed=4)[:, 0]
    a = OverSampledDFTAnalysisBankPtr(SampleFeaturePtr(x, M >> r, M >> r, True), h, M, m, r)
    with pytest.raises(jconsistency_error):
        a.current()                                   # frame index < 0 (stream.h:47-51)
    f0 = np.array(a.next(), copy=True)
    assert a.frameX() == 0 and a.size() == M and a.fftLen() == M
    assert np.array_equal(a.next(0), f0)              # same index: cached vector (stream.h:35-75)
    assert np.array_equal(a.current(), f0)
    X = bo.analysis(x, h, bo.BankGeometry(M, m, r, 0))
    n = 1
    while True:
        try:
            f = a.next()
        except StopIteration:
            break
        assert bo.rel_l2(f, X[n]) <= 1e-4
        n += 1
    assert n == X.shape[0] and a.isEnd()
    assert bo.rel_l2(f0, X[0]) <= 1e-4
    # spectra are Hermitian like the reference's full-M vectors
    assert np.allclose(f0[M // 2 + 1:], np.conj(f0[1:M // 2][::-1]))
    with pytest.raises(jconsistency_error):           # modulated.cc:269-271
        OverSampledDFTAnalysisBankPtr(SampleFeaturePtr(x, 128, 128, True), h[:-1], M, m, r)


def build_chain(bf, pcm, h, g, M, m, r):
    D = M >> r
    for c in range(pcm.shape[1]):
        sample = SampleFeaturePtr(pcm[:, c], blockLen=D, shiftLen=D, padZeros=True)
        bf.setChannel(OverSampledDFTAnalysisBankPtr(sample, h, M, m, r))
    return OverSampledDFTSynthesisBankPtr(bf, g, M, m, r)


def test_delay_and_sum_chain_is_fused_and_matches_oracle(prototypes):
    M, m, r, C, T = 256, 4, 1, 8, 24000
    h, g = proto(prototypes, M, m, r)
    mp = wl.circular_array(C)
    tau = wl.farfield_delays(mp, np.deg2rad(60), np.deg2rad(90))
    pcm = wl.array_recording(T, tau, seed=77)
    bf = SubbandDSPtr(fftLen=M)
    synth = build_chain(bf, pcm, h, g, M, m, r)
    with pytest.raises(j_error):                      # weights not computed yet (beamformer.cc:1140-1143)
        synth.next()
    with pytest.raises(jdimension_error):             # beamformer.cc:533-535
        bf.calcArrayManifoldVectors(FS, tau[:-1])
    bf.calcArrayManifoldVectors(FS, tau)
    out = np.concatenate([b.copy() for b in synth])
    assert synth.fused()
    geo = bo.BankGeometry(M, m, r, 0)
    _, Y, ref = bo.chain(pcm, h, g, geo, bo.ds_weights(tau, FS, M))
    assert bo.snr_db(out, ref) >= 70.0
    # the beamformer node on its own serves the same Y the oracle computes, as full-M Hermitian vectors
    bf.reset()
    y3 = [np.array(bf.next(), copy=True) for _ in range(4)][3]
    assert bo.rel_l2(y3, Y[3]) <= 1e-4
    assert bo.rel_l2(bf.getWeights(5), bo.ds_weights(tau, FS, M)[5]) <= 1e-12
    assert bf.chanN() == C and bf.dim() == M


def test_superdirective_chain_like_reference_driver(prototypes):
    M, m, r, C, T = 512, 2, 2, 16, 16000
    h, g = proto(prototypes, M, m, r)
    mp = wl.linear_array(C, 41.0)
    tau = wl.farfield_delays(mp, np.deg2rad(30), np.deg2rad(90))
    pcm = wl.array_recording(T, tau, seed=78)
    bf = SubbandMVDRPtr(fftLen=M)
    synth = build_chain(bf, pcm, h, g, M, m, r)
    bf.calcArrayManifoldVectors(FS, tau)
    assert bf.setDiffuseNoiseModel(mp, FS)
    assert not bf.setNoiseSpatialSpectralMatrix(3, np.eye(C + 1))      # wrong shape -> false (:2457-2464)
    bf.divideAllNonDiagonalElements(0.01)
    bf.setAllLevelsOfDiagonalLoading(0.1)
    assert bf.calcMVDRWeights(FS, 1.0e-8)
    out = np.concatenate([b.copy() for b in synth])
    assert synth.fused()
    Rn = bo.diagonal_load(bo.divide_nondiagonal(bo.diffuse_coherence(mp, FS, M), 0.01), 0.1)
    W = bo.mvdr_weights(Rn, bo.ds_weights(tau, FS, M))
    assert bo.rel_l2(np.stack([bf.getMVDRWeights(s) for s in range(M // 2 + 1)]), W) <= 1e-7
    _, _, ref = bo.chain(pcm, h, g, bo.BankGeometry(M, m, r, 0), W)
    assert bo.snr_db(out, ref) >= 70.0


def test_mvdr_node_refuses_to_serve_delay_and_sum_output(prototypes):
    """SubbandMVDR::next throws 'call calcMVDRWeights() once' when only the manifold exists (beamformer.cc:2587-2594), and a
    calcArrayManifoldVectors AFTER calcMVDRWeights leaves the MVDR weights in place (_wmvdr is separate from wq)."""
    M, m, r, C, T = 256, 4, 1, 4, 4000
    h, g = proto(prototypes, M, m, r)
    mp = wl.linear_array(C, 41.0)
    tau = wl.farfield_delays(mp, np.deg2rad(30), np.deg2rad(90))
    pcm = wl.array_recording(T, tau, seed=3)
    bf = SubbandMVDRPtr(fftLen=M)
    synth = build_chain(bf, pcm, h, g, M, m, r)
    with pytest.raises(j_error, match="calcArrayManifoldVectors"):
        bf.next()
    bf.calcArrayManifoldVectors(FS, tau)
    with pytest.raises(j_error, match="calcMVDRWeights"):
        bf.next()
    with pytest.raises(j_error, match="calcMVDRWeights"):
        next(iter(synth))
    assert bf.setDiffuseNoiseModel(mp, FS)
    bf.setAllLevelsOfDiagonalLoading(0.1)
    assert bf.calcMVDRWeights(FS, 1.0e-8)
    W = bo.mvdr_weights(bo.diagonal_load(bo.diffuse_coherence(mp, FS, M), 0.1), bo.ds_weights(tau, FS, M))
    bf.calcArrayManifoldVectors(FS, tau)               # must not replace the solved weights
    assert bo.rel_l2(np.stack([bf.getMVDRWeights(s) for s in range(M // 2 + 1)]), W) <= 1e-7
    out = np.concatenate([b.copy() for b in synth])
    _, _, ref = bo.chain(pcm, h, g, bo.BankGeometry(M, m, r, 0), W)
    assert bo.snr_db(out, ref) >= 70.0


def test_foreign_upstream_iterator_is_drained_frame_by_frame(prototypes):
    """A plain Python iterator in the middle of the chain (what PyVectorComplexFeatureStream allows,
    btk/stream/pyStream.h:89-130): the synthesis bank pulls its frames one by one."""
    M, m, r = 256, 4, 1
    h, g = proto(prototypes, M, m, r)
    x = wl.noise_recording(6000, 1, seed=5)[:, 0]
    geo = bo.BankGeometry(M, m, r, 0)
    X = bo.analysis(x, h, geo)
    synth = OverSampledDFTSynthesisBankPtr(iter(list(0.5 * X)), g, M, m, r)
    out = np.concatenate([b.copy() for b in synth])
    assert not synth.fused()
    assert bo.snr_db(out, 0.5 * bo.synthesis(X, g, geo).reshape(-1)) >= 70.0
