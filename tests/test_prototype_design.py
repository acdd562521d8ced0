"""de Haan prototype design (modulated/prototypeDesign.cc:223-272, 611-951).
CPU tier: numpy restatement against prototypes designed by the compiled reference (tests/golden/design_*.npz from
make_golden_design.py).  GPU tier: the device design (fp64 one-sided Jacobi pseudo-inverse) against the same fixtures,
against the oracle at L = 1024, and through the filter bank itself."""
import os

import numpy as np
import pytest

import btk_b200
import btk_oracle as bo
from conftest import GOLDEN

wl = btk_b200.workloads
CASES = sorted(f[len("design_"):-4] for f in os.listdir(GOLDEN)
               if f.startswith("design_") and f.endswith(".npz") and not f.startswith("design_nyquist_"))
# Nyquist(M)-constrained designs of the compiled reference (make_golden_nyquist.py): design_nyquist_<M>_<m>_<r>_<tol>.npz
NYQ = sorted(f[len("design_nyquist_"):-4] for f in os.listdir(GOLDEN) if f.startswith("design_nyquist_") and f.endswith(".npz"))


def _load(name):
    Z = np.load(os.path.join(GOLDEN, f"design_{name}.npz"))
    return {k: Z[k] for k in Z.files}


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_design(name):
    G = _load(name)
    M, m, r = [int(v) for v in G["geo"]]
    h = bo.design_analysis_dehaan(M, m, r, float(G["wp"]))
    assert _rel(h, G["h"]) <= 1e-9
    g = bo.design_synthesis_dehaan(G["h"], M, m, r, float(G["v"]))
    assert _rel(g, G["g"]) <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_design_matches_reference(name):
    G = _load(name)
    M, m, r = [int(v) for v in G["geo"]]
    h, eh = btk_b200.design_analysis_prototype(M, m, r, float(G["wp"]))
    assert _rel(h, G["h"]) <= 1e-8
    assert np.abs(eh - G["err_h"]).max() <= 1e-5            # dB
    g, eg = btk_b200.design_synthesis_prototype(G["h"], M, m, r, float(G["v"]), float(G["wp"]))
    assert _rel(g, G["g"]) <= 1e-8
    assert np.abs(eg - G["err_g"]).max() <= 1e-5
    with pytest.raises(btk_b200.BtkError):
        btk_b200.design_synthesis_prototype(G["h"][:-1], M, m, r)


@pytest.mark.gpu
def test_device_design_full_size_and_in_the_filter_bank():
    """(M, m, r) = (256, 4, 1), L = 1024 -- the geometry of BASELINE configs 1 and 2: the device design equals the float64
    oracle, and a bank built from the designed pair reconstructs a signal (gain 1/pi) with the fidelity the design's error figures promise."""
    M, m, r = 256, 4, 1
    h, eh = btk_b200.design_analysis_prototype(M, m, r)
    ho = bo.design_analysis_dehaan(M, m, r)
    assert _rel(h, ho) <= 1e-7
    g, eg = btk_b200.design_synthesis_prototype(h, M, m, r, 1.0)
    go = bo.design_synthesis_dehaan(h, M, m, r, 1.0)
    assert _rel(g, go) <= 1e-7
    nodes = btk_b200.streams
    d = nodes.AnalysisOversampledDFTDesignPtr(M, m, r, 1.0)
    assert np.array_equal(d.design(), h) and d.calcError(False).shape == (3,)
    # round trip through the device filter bank with the designed pair
    T, D, L = 40000, M >> r, M * m
    x = wl.noise_recording(T, 1, 7, sigma=2000.0)[:, 0].astype(np.float32)
    plan = btk_b200.Plan(M, m, r, 1, h, g)
    plan.set_weights(np.ones((plan.B, 1), dtype=np.complex128))
    y = plan.chain(x[:, None]).astype(np.float64)
    seg = x[3000:9000].astype(np.float64)
    scores = [float(np.dot(y[3000 + k: 9000 + k], seg)) for k in range(0, 2 * L + 1)]
    lag = int(np.argmax(scores))
    # (the chain skips its processing delay of 2m - 1 frames, so the residual lag is small: tau_h + tau_g - (2m-1) D - ...)
    n = T - lag - 2000
    a, b = y[lag + 1000: lag + 1000 + n], x[1000: 1000 + n].astype(np.float64)
    scale = float(np.dot(a, b) / np.dot(b, b))
    snr = 10 * np.log10(np.sum((scale * b) ** 2) / np.sum((a - scale * b) ** 2))
    # f = M/(pi D) h[2 tau - m] (prototypeDesign.cc:866) fixes the pair's overall gain at D/(pi D) = 1/pi
    assert abs(scale * np.pi - 1.0) < 0.05 and snr > 25.0, (scale, snr)
    plan.close()


# ------------------------------------------------------------------------------------------ Nyquist(M)-constrained designs
def _load_nyq(name):
    Z = np.load(os.path.join(GOLDEN, f"design_nyquist_{name}.npz"))
    return {k: Z[k] for k in Z.files}


@pytest.mark.parametrize("name", NYQ)
def test_oracle_matches_reference_nyquist_design(name):
    """AnalysisNyquistMDesign / SynthesisNyquistMDesign (prototypeDesign.cc:955-1119): numpy restatement against the
    compiled reference, both solution paths (the fixture with tolerance 0.2 takes "alternate solution 3" for g)."""
    G = _load_nyq(name)
    M, m, r = [int(v) for v in G["geo"]]
    tol = float(G["tol"])
    if M * m > 512:
        pytest.skip("L = 1024 restatement takes a minute of numpy SVDs: covered on the GPU tier")
    h, ph = bo.design_analysis_nyquist(M, m, r, 1.0, -1, tol, want_path=True)
    assert _rel(h, G["h"]) <= 1e-9 and ph == 4
    assert np.abs(h[::M] * M - (np.arange(m) == m // 2)).max() <= 1e-12         # the Nyquist(M) constraint itself
    g, pg = bo.design_synthesis_nyquist(G["h"], M, m, r, 1.0, -1, tol, want_path=True)
    assert _rel(g, G["g"]) <= 1e-9
    assert pg == (3 if tol > 0.1 else 4)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NYQ)
def test_device_nyquist_design_matches_reference(name):
    """The device designs (rectangular one-sided Jacobi decompositions, null-space projection) against the compiled
    reference's prototypes, (256, 4, 1) and (512, 2, 2) of the BASELINE geometries included: <= 1e-8 of the largest tap."""
    G = _load_nyq(name)
    M, m, r = [int(v) for v in G["geo"]]
    tol = float(G["tol"])
    h, ph = btk_b200.design_analysis_nyquist(M, m, r, 1.0, -1, tol)
    assert _rel(h, G["h"]) <= 1e-8 and ph == 4
    g, pg = btk_b200.design_synthesis_nyquist(G["h"], M, m, r, 1.0, -1, tol)
    assert _rel(g, G["g"]) <= 1e-8
    assert pg == (3 if tol > 0.1 else 4)
    d = btk_b200.streams.SynthesisNyquistMDesignPtr(G["h"], M, m, r, 1.0)
    assert np.array_equal(d.design(tol), g) and d.solutionPath() == pg
