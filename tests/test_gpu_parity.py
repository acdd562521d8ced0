"""GPU tier (-m gpu): the CUDA path, called through the C ABI (include/btkb200.h), against the oracle.

Gates (BASELINE.json north_star): relative L2 error <= 1e-4 on subband snapshots and beamformer outputs,
reconstructed time-domain SNR >= 70 dB against the reference output.  The oracle is float64; the device
computes in float32, so the typical figures are ~2e-7 and >120 dB.
"""
import numpy as np
import pytest

import btk_b200
import btk_oracle as bo
from conftest import golden_cases, load_golden, proto

pytestmark = pytest.mark.gpu
wl = btk_b200.workloads
TOL_REL = 1e-4      # north_star tolerance on snapshots / beamformer outputs
TOL_SNR = 70.0      # north_star tolerance on the reconstructed signal (dB)
FS = 16000.0


def half(X):
    """[F][C][M] full spectra -> [F][B][C] snapshots."""
    B = X.shape[2] // 2 + 1
    return X[:, :, :B].transpose(0, 2, 1)


@pytest.mark.parametrize("name", golden_cases())
def test_against_reference_golden(name, prototypes):
    """Committed outputs of the compiled reference: staged path and fused chain."""
    G = load_golden(name)
    M, m, r, dct, C, T = [int(v) for v in G["geo"]]
    h, g = prototypes[f"h_{M}_{m}_{r}"], prototypes[f"g_{M}_{m}_{r}"]
    plan = btk_b200.Plan(M, m, r, C, h, g, dct=dct)
    snap = plan.analysis(G["pcm"])
    assert snap.shape == (G["X"].shape[0], plan.B, C)
    assert bo.rel_l2(snap, half(G["X"])) <= TOL_REL
    if int(G["mode"]) == 1:
        plan.set_ds_weights(FS, G["delays"])
        plan.set_diffuse_noise_model(G["micpos"], FS)
        plan.diag_load(float(G["load"]))
        assert plan.solve_mvdr(FS, 1e-8) == 0
        # device double-precision solve vs the reference run with a double inverse ("oracle B")
        assert bo.rel_l2(plan.get_weights(), G["W"]) <= 1e-9
    else:
        plan.set_ds_weights(FS, G["delays"])
        assert bo.rel_l2(plan.get_weights(), G["W"]) <= 1e-12
    Y = plan.beamform(snap)
    assert bo.rel_l2(Y, G["Y"][:, : plan.B]) <= TOL_REL
    if "Y_floatsvd" in G:
        # the STOCK reference path (single-precision LINPACK SVD pseudoinverse, beamformer.cc:253-305): weights and
        # beamformer outputs stay inside the north_star tolerance of it as well
        assert bo.rel_l2(plan.get_weights(), G["W_floatsvd"]) <= TOL_REL
        assert bo.rel_l2(Y, G["Y_floatsvd"][:, : plan.B]) <= TOL_REL
    out = plan.synthesis(Y)
    assert out.shape == G["out"].shape
    assert bo.snr_db(out, G["out"]) >= TOL_SNR
    fused = plan.chain(G["pcm"])
    assert fused.shape == G["out"].shape
    assert bo.snr_db(fused, G["out"]) >= TOL_SNR
    if "S_cpp" in G:
        F = snap.shape[0]
        wts = (1 - 0.95) * 0.95 ** np.arange(F - 1, -1, -1.0)
        S = plan.covariance(snap, wts, conjugate=False)     # SpectralMatrixArray::update flavour (no conjugate)
        assert bo.rel_l2(S, G["S_cpp"][: plan.B]) <= TOL_REL
    plan.close()


CASES = [  # M, m, r, dct, C, T
    (256, 4, 1, 0, 1, 16000),     # BASELINE config 1 geometry (single channel round trip)
    (256, 4, 1, 0, 8, 40000),     # config 2 geometry (8-mic circular DS)
    (256, 4, 1, 1, 5, 9000),
    (256, 4, 1, 2, 2, 9000),
    (512, 2, 2, 0, 16, 20000),    # config 3 geometry
    (512, 2, 3, 0, 3, 7000),
    (512, 2, 2, 2, 7, 5000),
    (128, 2, 1, 0, 4, 6000),
    (64, 2, 1, 0, 6, 3000),
    (1024, 2, 2, 0, 4, 30000),
    (256, 2, 2, 0, 4, 8000),
    (512, 2, 1, 0, 4, 8000),
    (256, 4, 1, 0, 32, 6000),     # many channels: the launch keeps one CTA per SM (L2 footprint, capi.cu chain_prepare)
    (128, 2, 1, 0, 64, 4000),     # many channels, two CTAs per SM
]


@pytest.mark.parametrize("case", CASES)
def test_ds_chain_matches_oracle(case, prototypes):
    M, m, r, dct, C, T = case
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, dct)
    mp = wl.circular_array(C) if C > 1 else np.zeros((1, 3))
    tau = wl.farfield_delays(mp, np.deg2rad(60), np.deg2rad(90))
    pcm = wl.array_recording(T, tau, seed=20240 + M + C, noise_sigma=300.0)
    W = bo.ds_weights(tau, FS, M)
    X, Y, ref = bo.chain(pcm, h, g, geo, W)
    plan = btk_b200.Plan(M, m, r, C, h, g, dct=dct)
    plan.set_ds_weights(FS, tau)
    snap = plan.analysis(pcm)
    assert bo.rel_l2(snap, half(X)) <= TOL_REL
    Yd = plan.beamform(snap)
    assert bo.rel_l2(Yd, Y[:, : geo.B]) <= TOL_REL
    assert bo.snr_db(plan.synthesis(Yd), ref) >= TOL_SNR
    fused = plan.chain(pcm)
    assert bo.snr_db(fused, ref) >= TOL_SNR
    # first R-1 frames (priming quirk) and the tail frames individually
    D = geo.D
    for sl in (slice(0, geo.R * D), slice(-3 * D, None)):
        assert bo.snr_db(fused[sl], ref[sl]) >= TOL_SNR
    plan.close()


@pytest.mark.parametrize("C,M,m,r,load", [(16, 512, 2, 2, 1.0), (8, 256, 4, 1, 0.1), (64, 512, 2, 2, 1.0)])
def test_mvdr_chain_matches_oracle(C, M, m, r, load, prototypes):
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    mp = wl.linear_array(C, 41.0 if C <= 16 else 20.0)
    tau = wl.farfield_delays(mp, np.deg2rad(30), np.deg2rad(90))
    T = 12000
    pcm = wl.array_recording(T, tau, seed=31 + C, noise_sigma=200.0)
    wq = bo.ds_weights(tau, FS, M)
    Rn = bo.diagonal_load(bo.diffuse_coherence(mp, FS, M), load)
    W = bo.mvdr_weights(Rn, wq)
    X, Y, ref = bo.chain(pcm, h, g, geo, W)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    plan.set_ds_weights(FS, tau)
    plan.set_diffuse_noise_model(mp, FS)
    plan.diag_load(load)
    assert plan.solve_mvdr(FS, 1e-8) == 0
    assert bo.rel_l2(plan.get_weights(), W) <= 1e-7
    assert np.allclose(plan.get_covariance(3), Rn[3])
    Yd = plan.beamform(plan.analysis(pcm))
    assert bo.rel_l2(Yd, Y[:, : geo.B]) <= TOL_REL
    assert bo.snr_db(plan.chain(pcm), ref) >= TOL_SNR
    plan.close()


def test_sample_covariance_and_adaptive_mvdr(prototypes):
    """V1' (Python updateSx, Hermitian, ff = 0.99) -> diagonal loading -> solve -> apply."""
    M, m, r, C, T = 256, 4, 1, 8, 24000
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    mp = wl.circular_array(C)
    tau = wl.farfield_delays(mp, 1.0, 1.5)
    pcm = wl.noise_recording(T, C, seed=5, sigma=500.0)
    X = np.stack([bo.analysis(pcm[:, c], h, geo) for c in range(C)], axis=1)
    S = bo.spectral_matrix_py(X, 0.99)
    F = X.shape[0]
    wts = (1 - 0.99) * 0.99 ** np.arange(F - 1, -1, -1.0)
    wts[0] = 0.99 ** (F - 1)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    snap = plan.analysis(pcm)
    Sd = plan.covariance(snap, wts, conjugate=True)
    assert bo.rel_l2(Sd, S) <= TOL_REL
    plan.set_ds_weights(FS, tau)
    load = 1e-2 * float(np.real(np.trace(S[10]))) / C
    for s in range(geo.B):
        plan.set_covariance(s, Sd[s])
    plan.diag_load(load)
    assert plan.solve_mvdr() == 0
    W = bo.mvdr_weights(bo.diagonal_load(Sd, load), bo.ds_weights(tau, FS, M))
    assert bo.rel_l2(plan.get_weights(), W) <= 1e-6
    plan.close()


def test_adaptive_mvdr_on_device_end_to_end(prototypes):
    """cfg3 geometry (16-mic linear array, M=512 m=2 r=2): covariance estimated from a noise-only lead-in ON THE DEVICE
    (btkb200_estimate_covariance), loading 1e-2 trace/C, per-bin solve, then the fused chain on the recording."""
    M, m, r, C, T = 512, 2, 2, 16, 32000
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    mp = wl.linear_array(C, 41.0)
    tau = wl.farfield_delays(mp, np.deg2rad(30), np.deg2rad(90))
    noise = wl.noise_recording(16000, C, seed=41, sigma=300.0)
    pcm = wl.array_recording(T, tau, seed=42, noise_sigma=300.0)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    plan.set_ds_weights(FS, tau)
    for ff, last, conj in [(0.99, -1, True), (0.9, 20, True), (0.95, -1, False)]:
        plan.estimate_covariance(noise, forget=ff, last_frame=last, conjugate=conj)
        X = np.stack([bo.analysis(noise[:, c], h, geo) for c in range(C)], axis=1)
        if conj:
            S = bo.spectral_matrix_py(X[: (last + 1) if last >= 0 else None], ff)
        else:
            S = bo.spectral_matrix_cpp(X, ff)[: geo.B]
        Sd = np.stack([plan.get_covariance(s) for s in range(geo.B)])
        assert bo.rel_l2(Sd, S) <= TOL_REL
    plan.estimate_covariance(noise, forget=0.99)
    Sd = np.stack([plan.get_covariance(s) for s in range(geo.B)])
    load = 1e-2 * float(np.real(np.trace(Sd[40]))) / C
    plan.diag_load(load)
    assert plan.solve_mvdr() == 0
    W = bo.mvdr_weights(bo.diagonal_load(Sd, load), bo.ds_weights(tau, FS, M))
    assert bo.rel_l2(plan.get_weights(), W) <= 1e-6
    _, Y, ref = bo.chain(pcm, h, g, geo, W)
    assert bo.snr_db(plan.chain(pcm), ref) >= TOL_SNR
    # distortionless towards the look direction: w^H d = 1/C * C ... the reference normalisation gives w^H v = 1
    d = bo.ds_weights(tau, FS, M) * C
    assert np.allclose(np.sum(np.conj(W[1:]) * d[1:], axis=1), 1.0, atol=1e-6)
    plan.close()


def test_batch_ragged_and_edge_lengths(prototypes):
    M, m, r, C = 256, 4, 1, 4
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    tau = wl.farfield_delays(wl.circular_array(C), 0.3, 1.2)
    W = bo.ds_weights(tau, FS, M)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    plan.set_ds_weights(FS, tau)
    Ts = [1, 127, 128, 129, 1000, 5000, 12345, 256 * 40]   # shorter than a block, exact multiples, ragged
    pcms = [wl.noise_recording(T, C, seed=100 + i) for i, T in enumerate(Ts)]
    outs = plan.chain_batch(pcms)
    for pcm, out in zip(pcms, outs):
        _, _, ref = bo.chain(pcm, h, g, geo, W)
        assert out.shape == ref.shape
        assert bo.snr_db(out, ref) >= TOL_SNR
    # empty recording: no frames, no output
    assert plan.chain(np.zeros((0, C), np.float32)).size == 0
    plan.close()


def test_error_behaviour_matches_reference_contract(prototypes):
    M, m, r, C = 256, 4, 1, 4
    h, g = proto(prototypes, M, m, r)
    with pytest.raises(btk_b200.BtkError) as e:         # jconsistency_error, modulated.cc:269-271
        btk_b200.Plan(M, m, r, C, h[:-1], g)
    assert e.value.code == btk_b200._capi.EINVAL
    plan = btk_b200.Plan(M, m, r, C, h, g)
    pcm = wl.noise_recording(2000, C, 1)
    with pytest.raises(btk_b200.BtkError) as e:         # j_error "call calcArrayManifoldVectorsX() once", beamformer.cc:1140
        plan.chain(pcm)
    assert e.value.code == btk_b200._capi.ESTATE
    with pytest.raises(btk_b200.BtkError) as e:         # jdimension_error, beamformer.cc:533-535
        plan.set_ds_weights(FS, np.zeros(C + 1))
    assert e.value.code == btk_b200._capi.EINVAL
    with pytest.raises(btk_b200.BtkError) as e:         # calcMVDRWeights before a covariance, beamformer.cc:2394
        plan.set_ds_weights(FS, np.zeros(C))
        plan.solve_mvdr()
    assert e.value.code == btk_b200._capi.ESTATE
    with pytest.raises(btk_b200.BtkError):              # setNoiseSpatialSpectralMatrix shape check, :2457-2464
        plan.set_covariance(0, np.eye(C + 1))
    plan.close()


def test_roundtrip_property_full_size(prototypes):
    """BASELINE config 1 at full size through a size-independent property: analysis -> synthesis of a single
    channel reconstructs the (delayed, 1/D-scaled) input at the fidelity the shipped prototype allows
    (55.6 dB for (256,4,1)), and the GPU output equals the oracle's."""
    M, m, r = 256, 4, 1
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    T = 160000
    x = (wl.chirp(T) + wl.noise_recording(T, 1, 20241)[:, 0]).astype(np.float32)
    plan = btk_b200.Plan(M, m, r, 1, h, g)
    plan.set_weights(np.ones((geo.B, 1), dtype=np.complex128))
    y = plan.chain(x[:, None]) * geo.D
    # for the (256,4,1) Nyquist(M) pair the bank delay equals the synthesis priming (7 frames): lag 0 after the
    # start-up transient
    assert bo.snr_db(y[2048 : T - 4096], x[2048 : T - 4096]) > 50.0
    W = np.ones((geo.B, 1), dtype=np.complex128)
    _, _, ref = bo.chain(x[:, None], h, g, geo, W)
    assert bo.snr_db(y / geo.D, ref) >= TOL_SNR
    plan.close()


def test_linearity_property_full_size(prototypes):
    """Config 2 at full size (8 ch, 60 s): chain(a x1 + b x2) == a chain(x1) + b chain(x2)."""
    M, m, r, C, T = 256, 4, 1, 8, 960000
    h, g = proto(prototypes, M, m, r)
    tau = wl.farfield_delays(wl.circular_array(C), np.deg2rad(60), np.deg2rad(90))
    plan = btk_b200.Plan(M, m, r, C, h, g)
    plan.set_ds_weights(FS, tau)
    x1 = wl.array_recording(T, tau, seed=21240)
    x2 = wl.noise_recording(T, C, seed=21241)
    y1, y2, y12 = plan.chain(x1), plan.chain(x2), plan.chain((0.5 * x1 - 2.0 * x2).astype(np.float32))
    assert bo.snr_db(y12, 0.5 * y1.astype(np.float64) - 2.0 * y2.astype(np.float64)) > 90.0
    # steering at the true direction of arrival passes the source: output power ~ source power / D^2
    assert np.std(y1) * 128 > 0.5 * 8000 / np.sqrt(2)
    plan.close()


def test_dct2_with_unit_prototype_span_is_refused(prototypes):
    """delayCompensationType 2 with m*R = 1: the reference's look-ahead m R / 2 - 1 underflows (modulated.cc:285-290)."""
    h, g = wl.kaiser_prototype(256, 1, 0)
    with pytest.raises(btk_b200.BtkError) as e:
        btk_b200.Plan(256, 1, 0, 2, h, g, dct=2)
    assert e.value.code == btk_b200._capi.EINVAL


def test_large_batch_chunk_model_matches_single_recording_calls(prototypes):
    """A batch big enough for choose_chunk_model (more than two waves of CTAs) against the same recordings one by one
    (small-job chunking, itself checked against the oracle above): the chunk size must not change the result beyond
    float32 rounding (a frame can be packed with a different partner frame)."""
    M, m, r, C = 256, 4, 1, 2
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    tau = wl.farfield_delays(wl.circular_array(C), 0.4, 1.3)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    plan.set_ds_weights(FS, tau)
    rng = np.random.default_rng(11)
    pcms = [wl.noise_recording(int(rng.integers(50000, 70000)), C, seed=500 + i, sigma=900.0) for i in range(64)]
    outs = plan.chain_batch(pcms)
    for i in (0, 1, 17, 40, 63):
        single = plan.chain(pcms[i])
        assert outs[i].shape == single.shape and bo.snr_db(outs[i], single) >= 100.0
    W = bo.ds_weights(tau, FS, M)
    assert bo.snr_db(outs[5], bo.chain(pcms[5], h, g, geo, W)[2]) >= TOL_SNR
    plan.close()


def test_chain_batch_multi_plans(prototypes):
    """btkb200_chain_batch_multi (the C ABI's multi-GPU entry): recordings round-robin over the plans, one host
    thread per plan, no inter-GPU traffic.  One plan per device on every visible device (two plans on device 0 when the
    box has a single GPU, which still drives the threaded path); ragged lengths; each output against the oracle and
    against the single-plan batch call."""
    M, m, r, C = 256, 4, 1, 4
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    tau = wl.farfield_delays(wl.linear_array(C, 41.0), np.deg2rad(30), np.deg2rad(90))
    ndev = btk_b200.device_count()
    devs = list(range(ndev)) if ndev >= 2 else [0, 0]
    plans = [btk_b200.Plan(M, m, r, C, h, g, device=d) for d in devs]
    for p in plans:
        p.set_ds_weights(FS, tau)
    Ts = [5000, 8000, 3001, 6400, 129, 7000, 4096]
    xs = [np.ascontiguousarray(wl.array_recording(T, tau, seed=40 + i)) for i, T in enumerate(Ts)]
    outs = [np.zeros(geo.nblk(T) * geo.D, np.float32) for T in Ts]
    btk_b200._capi.chain_batch_multi(plans, xs, outs)
    single = plans[0].chain_batch(xs)
    W = bo.ds_weights(tau, FS, M)
    for i, x in enumerate(xs):
        assert np.array_equal(outs[i], single[i])                 # the same kernel on the same data, whatever the device
        assert bo.snr_db(outs[i], bo.chain(x, h, g, geo, W)[2]) >= TOL_SNR
    # ownership is the round-robin rule of btk_b200.sharding
    for d in range(len(plans)):
        assert btk_b200.sharding.shard(len(xs), d, len(plans)) == [i for i in range(len(xs)) if i % len(plans) == d]
    for p in plans:
        p.close()
