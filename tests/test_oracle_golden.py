"""CPU tier: the numpy oracle (oracle/btk_oracle.py) against the golden vectors produced by the compiled
reference (tests/golden/make_golden.py), and -- when oracle/_ref/libbtk_ref.so is present -- against the
reference itself on fresh seeds.  This is what pins the oracle (the reference ships no known-answer tests)."""
import numpy as np
import pytest

import btk_oracle as bo
from conftest import golden_cases, load_golden


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference_golden(name, prototypes):
    G = load_golden(name)
    M, m, r, dct, C, T = [int(v) for v in G["geo"]]
    h, g = prototypes[f"h_{M}_{m}_{r}"], prototypes[f"g_{M}_{m}_{r}"]
    geo = bo.BankGeometry(M, m, r, dct)
    pcm = G["pcm"]
    assert G["X"].shape == (geo.analysis_frames(T), C, M)
    X = np.stack([bo.analysis(pcm[:, c], h, geo) for c in range(C)], axis=1)
    assert bo.rel_l2(X, G["X"]) < 1e-13
    wq = bo.ds_weights(G["delays"], 16000.0, M)
    if int(G["mode"]) == 1:
        Rn = bo.diagonal_load(bo.diffuse_coherence(G["micpos"], 16000.0, M), float(G["load"]))
        W = bo.mvdr_weights(Rn, wq)
    else:
        W = wq
    assert bo.rel_l2(W, G["W"]) < 1e-11
    Y = bo.beamform(X, W)
    assert bo.rel_l2(Y, G["Y"]) < 1e-11
    out = bo.synthesis(Y, g, geo).reshape(-1)
    assert out.shape == G["out"].shape
    assert bo.snr_db(out, G["out"]) > 140.0
    if "S_cpp" in G:
        assert bo.rel_l2(bo.spectral_matrix_cpp(X, 0.95), G["S_cpp"]) < 1e-12


def test_frame_counts_match_survey(prototypes):
    # 132/128/253 analysis frames at T=16000 (dct 0), 129 with dct=2; 125/125/250 synthesis frames
    for (M, m, r), fa, fs in [((256, 4, 1), 132, 125), ((512, 2, 2), 128, 125), ((512, 2, 3), 253, 250)]:
        geo = bo.BankGeometry(M, m, r, 0)
        assert geo.analysis_frames(16000) == fa
        assert geo.synthesis_frames(fa) == fs
    assert bo.BankGeometry(256, 4, 1, 2).analysis_frames(16000) == 129


@pytest.mark.skipif(not bo.CompiledReference.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("cfg", [(256, 4, 1, 0), (256, 4, 1, 2), (512, 2, 2, 1), (512, 2, 3, 0)])
def test_oracle_matches_compiled_reference(cfg, prototypes):
    M, m, r, dct = cfg
    ref = bo.CompiledReference()
    h, g = prototypes[f"h_{M}_{m}_{r}"], prototypes[f"g_{M}_{m}_{r}"]
    geo = bo.BankGeometry(M, m, r, dct)
    rng = np.random.default_rng(99)
    C, T = 3, 2500
    pcm = (1000 * rng.standard_normal((T, C))).astype(np.float32)
    tau = np.array([0.0, 1.1e-4, -0.7e-4])
    res = ref.chain(pcm, h, g, geo, tau)
    W = bo.ds_weights(tau, 16000.0, M)
    X, Y, out = bo.chain(pcm, h, g, geo, W)
    assert bo.rel_l2(X, res["X"]) < 1e-13
    assert bo.rel_l2(Y, res["Y"]) < 1e-13
    assert bo.snr_db(out, res["out"]) > 140.0
    # error behaviour the drop-in classes must reproduce: JCONSISTENCY=3, JDIMENSION=4, JERROR=0, JITERATOR=8
    # (common/jexception.h:41-57; the SWIG layer maps code 8 to StopIteration, include/jexception.i:63-69)
    assert [ref.error_probe(i) for i in range(4)] == [3, 4, 0, 8]
