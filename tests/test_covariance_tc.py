"""GPU tier: the tensor-core covariance kernel (csrc/kern_cov_tc.cu: tcgen05.mma kind::tf32 with hi/lo split inputs,
FP32 accumulation in TMEM) against a float64 restatement of SpectralMatrixArray::update (beamformer.cc:142-163, x x^T)
and SubbandBeamformerMVDR.updateSx (lib/subbandBeamforming.py:1170-1175, x x^H) with the recursions unrolled into
per-frame weights.  Gate: rel-L2 <= 1e-4 (north_star); the split keeps it near 1e-6."""
import numpy as np
import pytest

import btk_b200
import btk_oracle as bo

pytestmark = pytest.mark.gpu
wl = btk_b200.workloads


def _gram(X, wts, conj):
    # X [F][B][C] complex128
    Y = np.conj(X) if conj else X
    return np.einsum("f,fba,fbc->bac", wts, X, Y)


@pytest.mark.parametrize("C,F,conj", [(64, 700, True), (64, 33, False), (24, 1, True), (40, 513, True), (33, 96, False),
                                      (64, 1300, True), (16, 1253, True), (16, 100, False), (8, 333, True), (5, 64, True),
                                      (3, 40, False), (1, 77, True), (12, 90, True)])
def test_tensor_core_covariance_matches_float64(C, F, conj):
    M, m, r = 64, 2, 1
    h, g = wl.kaiser_prototype(M, m, r)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    rng = np.random.default_rng(C * 1000 + F)
    B = plan.B
    X = (rng.standard_normal((F, B, C)) + 1j * rng.standard_normal((F, B, C))) * rng.uniform(10.0, 3000.0, (1, B, 1))
    X = X.astype(np.complex64)
    ff = 0.99
    wts = (1 - ff) * ff ** np.arange(F - 1, -1, -1.0)
    wts[0] = ff ** (F - 1)
    S = plan.covariance(X, wts, conjugate=conj)
    ref = _gram(X.astype(np.complex128), wts, conj)
    assert S.shape == ref.shape
    err = bo.rel_l2(S, ref)
    assert err <= 1e-4, err
    assert err <= 5e-6, f"hi/lo split lost precision: {err}"
    # per-bin check as well (a bin with a small scale must not drown in the others)
    per_bin = max(bo.rel_l2(S[b], ref[b]) for b in range(B))
    assert per_bin <= 1e-5, per_bin
    if conj:
        assert np.allclose(S, np.conj(S.transpose(0, 2, 1)), rtol=0, atol=1e-6 * np.abs(ref).max())
    plan.close()
