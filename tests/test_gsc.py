"""SubbandGSC with fixed active weights (beamformer/beamformer.cc:398-479 blocking matrix, :761-783
sidelobe canceller, :1251-1356 output).  CPU tier: numpy restatement against outputs of the compiled reference
(tests/golden/gsc_*.npz from make_golden_gsc.py).  GPU tier: the C-ABI setup (host, double) + the device apply / fused
chain against the same fixtures."""
import os

import numpy as np
import pytest

import btk_b200
import btk_oracle as bo
from conftest import GOLDEN, proto

FS = 16000.0
CASES = sorted(f[len("gsc_"):-4] for f in os.listdir(GOLDEN) if f.startswith("gsc_") and f.endswith(".npz"))


def _load(name):
    Z = np.load(os.path.join(GOLDEN, f"gsc_{name}.npz"))
    return {k: Z[k] for k in Z.files}


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_outputs(name, prototypes):
    G = _load(name)
    M, m, r, dct, C, T = [int(v) for v in G["geo"]]
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, dct)
    wq = bo.ds_weights(G["delays"], FS, M)
    assert np.abs(wq - G["wq"]).max() <= 1e-14
    for s in (0, 1, geo.B // 2, geo.B - 1):
        Bm = bo.blocking_matrix(wq[s])
        assert np.abs(Bm - G["Bm"][s]).max() <= 1e-12
    W = bo.gsc_weights(wq, G["wa"], bool(G["normalize"]))
    _, Y, out = bo.chain(G["pcm"], h, g, geo, W)
    assert bo.rel_l2(Y, G["Y"]) <= 1e-12
    assert bo.snr_db(out, G["out"]) >= 120.0


def test_blocking_matrix_properties():
    """Columns are orthonormal; the projector is I - conj(v) v^T/|v|^2 (the reference's zgeru is unconjugated), so the
    columns are orthogonal to v in the bilinear sense v^T b = 0 -- the property SubbandGSC relies on is kept as is."""
    rng = np.random.default_rng(3)
    v = rng.standard_normal(7) + 1j * rng.standard_normal(7)
    Bm = bo.blocking_matrix(v)
    assert Bm.shape == (7, 6)
    assert np.abs(np.conj(Bm).T @ Bm - np.eye(6)).max() <= 1e-12
    assert np.abs(v @ Bm).max() <= 1e-12
    with pytest.raises(ValueError):
        bo.blocking_matrix(np.ones(1))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_gsc_matches_reference_outputs(name, prototypes):
    G = _load(name)
    M, m, r, dct, C, T = [int(v) for v in G["geo"]]
    h, g = proto(prototypes, M, m, r)
    plan = btk_b200.Plan(M, m, r, C, h, g, dct=dct)
    with pytest.raises(btk_b200.BtkError) as e:           # setActiveWeights_f before calcGSCWeights: j_error (:1427-1430)
        plan.gsc_set_active_weights(1, np.zeros(2 * (C - 1)))
    assert e.value.code == btk_b200._capi.ESTATE
    plan.gsc_calc_weights(FS, G["delays"])
    with pytest.raises(btk_b200.BtkError) as e:           # jdimension_error (:764-766)
        plan.gsc_set_active_weights(1, np.zeros(2 * C))
    assert e.value.code == btk_b200._capi.EINVAL
    B = plan.B
    for s in (0, 1, B - 1):
        assert np.abs(plan.gsc_blocking_matrix(s) - G["Bm"][s]).max() <= 1e-12
    for s in range(B):
        plan.gsc_set_active_weights(s, G["wa"][s].view(np.float64))
    plan.gsc_apply(bool(G["normalize"]))
    W = bo.gsc_weights(G["wq"], G["wa"], bool(G["normalize"]))
    assert bo.rel_l2(plan.get_weights(), W) <= 1e-12
    Y = plan.beamform(plan.analysis(G["pcm"]))
    assert bo.rel_l2(Y, G["Y"][:, :B]) <= 1e-4
    out = plan.chain(G["pcm"])
    assert out.shape == G["out"].shape and bo.snr_db(out, G["out"]) >= 70.0
    # zeroActiveWeights: back to the quiescent (delay-and-sum) beamformer
    plan.gsc_zero_active_weights()
    plan.gsc_apply(False)
    assert bo.rel_l2(plan.get_weights(), G["wq"]) <= 1e-14
    plan.close()


@pytest.mark.gpu
def test_gsc_single_channel_rejected(prototypes):
    h, g = proto(prototypes, 256, 4, 1)
    plan = btk_b200.Plan(256, 4, 1, 1, h, g)
    with pytest.raises(btk_b200.BtkError) as e:           # jdimension_error, beamformer.cc:536-539
        plan.gsc_calc_weights(FS, np.zeros(1))
    assert e.value.code == btk_b200._capi.EINVAL
    plan.close()


@pytest.mark.gpu
def test_gsc_stream_node(prototypes):
    """banks -> SubbandGSC (calcGSCWeights, setActiveWeights_f, normalizeWeight) -> synthesis through the drop-in nodes;
    the fused chain runs with the effective weights."""
    G = _load(CASES[0])
    M, m, r, dct, C, T = [int(v) for v in G["geo"]]
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, dct)
    bf = btk_b200.SubbandGSCPtr(M, False)
    for c in range(C):
        src = btk_b200.SampleFeaturePtr(G["pcm"][:, c], blockLen=geo.D, shiftLen=geo.D, padZeros=True)
        bf.setChannel(btk_b200.OverSampledDFTAnalysisBankPtr(src, h, M, m, r, dct))
    with pytest.raises(btk_b200.streams.j_error):
        bf.next()
    bf.calcGSCWeights(FS, G["delays"])
    for s in range(M):
        bf.setActiveWeights_f(s, G["wa"][min(s, M - s)].view(np.float64))
    assert np.abs(bf.getBlockingMatrix(0, 3) - G["Bm"][3]).max() <= 1e-12
    syn = btk_b200.OverSampledDFTSynthesisBankPtr(bf, g, M, m, r, dct)
    out = np.concatenate([np.array(f, copy=True) for f in syn])
    assert syn.fused()
    assert out.shape == G["out"].shape and bo.snr_db(out, G["out"]) >= 70.0
