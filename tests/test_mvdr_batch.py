"""Batched, device-resident adaptive MVDR (btkb200_mvdr_chain_batch): per recording covariance -> diagonal loading ->
per-bin solve -> fused chain with that recording's weights.  Checked against (a) the same steps through the single-
recording entry points (host round trips between them) and (b) the float64 oracle (SpectralMatrixArray / updateSx
recursion -> loading -> MVDR weights -> chain).  The loadings are chosen around 1e-2 of the per-channel power: with a
negligible load and a short adaptation window the matrices have condition numbers of 1e7 and above, and the float32
rounding of the snapshots alone moves the weights by more than the comparison tolerance (in every implementation)."""
import numpy as np
import pytest

import btk_b200
import btk_oracle as bo
from conftest import proto

pytestmark = pytest.mark.gpu
wl = btk_b200.workloads
FS = 16000.0


def _oracle_out(pcm, h, g, geo, tau, forget, last, conj, load_abs, load_rel):
    C = pcm.shape[1]
    X = np.stack([bo.analysis(pcm[:, c], h, geo) for c in range(C)], axis=1)
    Xa = X[: (last + 1) if last >= 0 else None]
    S = bo.spectral_matrix_py(Xa, forget) if conj else bo.spectral_matrix_cpp(Xa, forget)[: geo.B]
    Rl = np.array(S[: geo.B], dtype=np.complex128)
    for s in range(geo.B):
        add = float(np.float32(load_abs)) + load_rel * float(np.real(np.trace(Rl[s]))) / C
        Rl[s] = Rl[s] + add * np.eye(C)
    W = bo.mvdr_weights(Rl, bo.ds_weights(tau, FS, geo.M))
    return bo.chain(pcm, h, g, geo, W)[2], W


@pytest.mark.parametrize("cfg", [(256, 4, 1, 8, True, 0.99, -1, 0.0, 1e-2), (512, 2, 2, 16, True, 0.95, 30, 5.0e5, 0.0),
                                 (512, 2, 2, 6, False, 0.95, -1, 5.0e6, 0.0), (256, 4, 1, 3, True, 0.9, 0, 10.0, 0.1)])
def test_batch_matches_single_recording_path_and_oracle(cfg, prototypes):
    M, m, r, C, conj, forget, last, load_abs, load_rel = cfg
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    mp = wl.linear_array(C, 41.0)
    tau = wl.farfield_delays(mp, np.deg2rad(30), np.deg2rad(90))
    Ts = [9000, 4000, 12345]
    pcms = [wl.array_recording(T, tau, seed=60 + i, noise_sigma=300.0 + 100 * i) for i, T in enumerate(Ts)]
    plan = btk_b200.Plan(M, m, r, C, h, g)
    plan.set_ds_weights(FS, tau)
    outs, nfb = plan.mvdr_chain_batch(pcms, forget=forget, last_frame=last, conjugate=conj, load_abs=load_abs, load_rel=load_rel)
    assert nfb.tolist() == [0, 0, 0]
    for pcm, out in zip(pcms, outs):
        # (a) the same pipeline through the single-recording entry points
        plan.estimate_covariance(pcm, forget=forget, last_frame=last, conjugate=conj)
        for s in range(geo.B):
            tr = float(np.real(np.trace(plan.get_covariance(s))))
            if load_abs:
                plan.diag_load(load_abs, s)
            if load_rel:
                plan.diag_load(load_rel * tr / C, s)
        assert plan.solve_mvdr() == 0
        ref_dev = plan.chain(pcm)
        assert out.shape == ref_dev.shape and bo.snr_db(out, ref_dev) >= 80.0
        # (b) float64 oracle
        ref, W = _oracle_out(pcm, h, g, geo, tau, forget, last, conj, load_abs, load_rel)
        assert bo.snr_db(out, ref) >= 70.0
    # the plan's own weights are untouched by the batch call (still what solve_mvdr installed last)
    plan.close()


def test_batch_requires_manifold_and_limits(prototypes):
    h, g = proto(prototypes, 256, 4, 1)
    plan = btk_b200.Plan(256, 4, 1, 4, h, g)
    x = [wl.noise_recording(3000, 4, seed=1)]
    with pytest.raises(btk_b200.BtkError) as e:
        plan.mvdr_chain_batch(x)
    assert e.value.code == btk_b200._capi.ESTATE
    plan.set_ds_weights(FS, np.zeros(4))
    with pytest.raises(btk_b200.BtkError) as e:
        plan.mvdr_chain_batch(x, forget=1.5)
    assert e.value.code == btk_b200._capi.EINVAL
    outs, nfb = plan.mvdr_chain_batch([])
    assert outs == [] and nfb.size == 0
    plan.close()
