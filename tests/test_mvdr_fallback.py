"""V3: the identity fallback of SubbandMVDR::calcMVDRWeights (beamformer.cc:2425-2427), taken when
pseudoinverse() drops a singular value below dThreshold (:275-283) -- against the STOCK compiled reference (single-precision
LINPACK SVD, "oracle A").

The device rule: smallest singular value (estimated from the triangular factor by inverse iteration) below dThreshold, or a
vanishing pivot.  Matrices are built with singular values far on either side of the threshold, where the reference's answer
does not depend on the rounding of its float SVD: sigma_max * 6e-8 (its noise floor) stays below 1e-8 for the rejected
bins, the accepted bins are well conditioned.
"""
import numpy as np
import pytest

import btk_b200
import btk_oracle as bo
from conftest import proto

wl = btk_b200.workloads
FS = 16000.0
pytestmark = pytest.mark.gpu


def hermitian(rng, C, sv):
    Q, _ = np.linalg.qr(rng.normal(size=(C, C)) + 1j * rng.normal(size=(C, C)))
    return (Q * np.asarray(sv)) @ Q.conj().T


@pytest.mark.skipif(not bo.CompiledReference.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("C,M,m,r", [(6, 256, 4, 1), (16, 512, 2, 2)])
def test_identity_fallback_matches_stock_reference(prototypes, C, M, m, r):
    ref = bo.CompiledReference()
    rng = np.random.default_rng(C)
    B = M // 2 + 1
    geo = bo.BankGeometry(M, m, r, 0)
    mp = wl.linear_array(C, 41.0)
    tau = wl.farfield_delays(mp, np.deg2rad(30), np.deg2rad(90))
    Rn = np.zeros((B, C, C), np.complex128)
    kind = np.zeros(B, int)
    for s in range(B):
        k = s % 5
        kind[s] = k
        if k == 0:
            Rn[s] = hermitian(rng, C, rng.uniform(0.5, 2.0, C))                           # well conditioned
        elif k == 1:
            Rn[s] = hermitian(rng, C, np.r_[np.full(C - 1, 0.02), 1e-12])                 # one singular value far below 1e-8
        elif k == 2:
            Rn[s] = np.diag(np.r_[np.ones(C - 1), 0.0]).astype(np.complex128)             # exactly singular
        elif k == 3:
            Rn[s] = np.zeros((C, C))                                                      # nothing to invert
        else:
            Rn[s] = hermitian(rng, C, np.r_[np.full(C - 2, 0.02), 3e-11, 1e-13]) * 0.5    # rank C - 2 at the threshold's scale
    h, g = proto(prototypes, M, m, r)
    pcm = wl.array_recording(4000, tau, seed=9, noise_sigma=500.0)
    R = ref.chain(pcm, h, g, geo, tau, mode="mvdr", Rn=Rn, dThreshold=1e-8, inverse="float", want_snap=False)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    plan.set_ds_weights(FS, tau)
    for s in range(B):
        plan.set_covariance(s, Rn[s])
    nfb = plan.solve_mvdr(FS, 1e-8)
    w = plan.get_weights()
    wq = bo.ds_weights(tau, FS, M)
    ident = wq / (np.sum(np.abs(wq) ** 2, axis=1, keepdims=True) * C)        # t = d: w = d / (d^H d C)
    rejected = kind != 0
    rejected[0] = False                                                       # bin 0 is never solved (:2410-2415)
    assert nfb == int(rejected.sum())
    # the stock reference took the same branch in every bin ...
    assert np.abs(R["W"][rejected] - ident[rejected]).max() <= 1e-12
    # ... and so did the device
    assert np.abs(w[rejected] - ident[rejected]).max() <= 1e-12
    ok = ~rejected
    ok[0] = False
    assert bo.rel_l2(w[ok], R["W"][ok]) <= 1e-5                                # float SVD against the double solve
    assert np.abs(w[0] - 1.0).max() == 0.0 and np.abs(R["W"][0] - 1.0).max() == 0.0
    # the chain with those weights
    out = plan.chain(pcm)
    assert out.shape == R["out"].shape and bo.snr_db(out, R["out"]) >= 70.0
    plan.close()


def test_threshold_is_on_singular_values_not_pivots(prototypes):
    """a near-singular R whose pivots all stay above the threshold.  A = L U with unit pivots hides
    sigma_min: the classic bidiagonal example U = I - 2 * superdiagonal has sigma_min ~ 2^-(C-1)."""
    C, M, m, r = 40, 256, 4, 1
    B = M // 2 + 1
    U = np.eye(C) - 2.0 * np.eye(C, k=1)
    A = (U.conj().T @ U).astype(np.complex128)                # Hermitian, every pivot of the elimination is O(1)
    smin = np.linalg.svd(A, compute_uv=False)[-1]
    assert smin < 1e-12                                        # true value 2^-78; numpy shows its own rounding floor
    h, g = proto(prototypes, M, m, r)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    tau = wl.farfield_delays(wl.linear_array(C, 20.0), np.deg2rad(30), np.deg2rad(90))
    plan.set_ds_weights(FS, tau)
    good = np.eye(C, dtype=np.complex128)
    for s in range(B):
        plan.set_covariance(s, A if s % 2 else good)
    nfb = plan.solve_mvdr(FS, 1e-8)
    assert nfb == B // 2                                       # every odd bin rejected, although no pivot is small
    assert plan.solve_mvdr(FS, 0.0) == 0                       # threshold 0: nothing is rejected (only a vanishing pivot would be)
    plan.close()
