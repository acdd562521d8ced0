"""GPU tier: the warp-specialised fused chain (csrc/chain_ws.cuh) -- producer warpgroup + mbarrier stages, and the
channel split over a thread-block cluster with the partial beamformer outputs summed through distributed shared
memory -- against the oracle, for every cluster size a shape admits, and against the first sessions' kernel.

Reference behaviour checked: modulated/modulated.cc:412-516, 595-664 and beamformer/beamformer.cc:1137-1200 (the fused
analysis -> weight apply -> synthesis chain); gates as everywhere: >= 70 dB SNR on the reconstructed PCM."""
import numpy as np
import pytest

import btk_b200
import btk_oracle as bo
from conftest import proto

pytestmark = pytest.mark.gpu
wl = btk_b200.workloads
FS = 16000.0
EUNSUPPORTED = btk_b200._capi.EUNSUPPORTED


def _weights(rng, geo, C, M):
    mp = wl.linear_array(C, 20.0)
    tau = wl.farfield_delays(mp, np.deg2rad(30), np.deg2rad(90))
    return bo.ds_weights(tau, FS, M) * np.exp(1j * rng.uniform(0, 2 * np.pi, (geo.B, C)))


@pytest.mark.parametrize("shape", [(256, 4, 1), (512, 2, 2), (512, 2, 3), (128, 2, 1), (1024, 2, 1), (256, 2, 0)])
@pytest.mark.parametrize("C", [4, 8, 16, 24, 64])
def test_ws_chain_every_cluster_size(shape, C, prototypes):
    M, m, r = shape
    D = M >> r
    rng = np.random.default_rng(M + 7 * C + r)
    try:
        h, g = proto(prototypes, M, m, r)
    except Exception:
        h, g = wl.designed_prototype(M, m, r) if M * m <= 2048 else wl.kaiser_prototype(M, m, r)   # designed on the device
    geo = bo.BankGeometry(M, m, r, 0)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    W = _weights(rng, geo, C, M)
    plan.set_weights(W)
    Ts = [int(rng.integers(70 * D, 90 * D)), int(rng.integers(3 * D, 9 * D)) + 5]
    pcms = [wl.noise_recording(T, C, seed=int(rng.integers(1 << 30)), sigma=700.0) for T in Ts]
    refs = [bo.chain(x, h, g, geo, W)[2] for x in pcms]
    try:
        plan.tune(chain_ws=1)
    except btk_b200.BtkError as e:
        assert e.code == EUNSUPPORTED
        pytest.skip(f"no warp-specialised kernel for {shape}")
    ran = []
    for S in (1, 2, 4, 8):
        try:
            plan.tune(cluster=S)
        except btk_b200.BtkError as e:
            assert e.code == EUNSUPPORTED       # S does not divide the channel groups / the frame pairs
            continue
        outs = plan.chain_batch(pcms)
        assert plan.tuning() == {"chain_ws": 1, "cluster": S}
        for out, ref in zip(outs, refs):
            assert out.shape == ref.shape
            assert bo.snr_db(out, ref) >= 70.0, f"cluster {S}"
        ran.append(S)
    assert 1 in ran
    # the automatic choice, and the first sessions' kernel on the same input
    plan.tune(cluster=0)
    auto = plan.chain_batch(pcms)
    plan.tune(chain_ws=0)
    old = plan.chain_batch(pcms)
    assert plan.tuning()["chain_ws"] == 0
    for a, o, ref in zip(auto, old, refs):
        assert bo.snr_db(a, ref) >= 70.0 and bo.snr_db(o, ref) >= 70.0
    plan.close()


def test_ws_chain_long_ragged_batch(prototypes):
    """Many work items per launch: several waves of CTAs, chunk boundaries inside recordings, per-recording tables."""
    M, m, r, C = 256, 4, 1, 8
    D = M >> r
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    rng = np.random.default_rng(3)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    W = _weights(rng, geo, C, M)
    plan.set_weights(W)
    Ts = [int(v) for v in rng.integers(200 * D, 1500 * D, 12)] + [1, D - 1, D, D + 1]
    pcms = [wl.noise_recording(T, C, seed=100 + i, sigma=500.0) for i, T in enumerate(Ts)]
    outs = plan.chain_batch(pcms)
    assert plan.tuning()["chain_ws"] == 1
    for x, out in zip(pcms, outs):
        ref = bo.chain(x, h, g, geo, W)[2]
        assert out.shape == ref.shape and bo.snr_db(out, ref) >= 70.0
    # idempotent: the same call again gives the same bits (no stale stage, no barrier-phase carry-over)
    again = plan.chain_batch(pcms)
    for a, b in zip(outs, again):
        assert np.array_equal(a, b)
    plan.close()


@pytest.mark.parametrize("cfg", [(256, 4, 1, 8), (512, 2, 2, 16), (128, 2, 1, 4), (256, 2, 1, 12), (128, 4, 1, 8)])
def test_ws_chain_many_short_recordings(cfg, prototypes):
    """Persistent schedule + overlap-add warpgroup (chain_ws.cuh 'WsSegs', 'chain_ws_synth_iter'): hundreds of short recordings
    make every CTA walk several segments -- history ring zeroed per segment, stage counters and the tensor-memory slots
    running on across them, transform warps an iteration ahead of the overlap-add warps -- and a few long ones are cut
    between CTAs.  Every output against the oracle, and the same bits on a second run."""
    M, m, r, C = cfg
    D = M >> r
    try:
        h, g = proto(prototypes, M, m, r)
    except Exception:
        h, g = wl.kaiser_prototype(M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    rng = np.random.default_rng(M + C)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    W = _weights(rng, geo, C, M)
    plan.set_weights(W)
    Ts = [int(v) for v in rng.integers(1, 40 * D, 300)] + [0, 0, 5000 * D + 3, 1, 2500 * D]
    rng.shuffle(Ts)
    pcms = [wl.noise_recording(T, C, seed=1000 + i, sigma=500.0) for i, T in enumerate(Ts)]
    outs = plan.chain_batch(pcms)
    assert plan.tuning()["chain_ws"] == 1
    for i in list(range(0, len(Ts), 7)) + [Ts.index(5000 * D + 3), Ts.index(2500 * D)]:
        x, out = pcms[i], outs[i]
        if x.shape[0] == 0:
            assert out.size == 0
            continue
        ref = bo.chain(x, h, g, geo, W)[2]
        assert out.shape == ref.shape and bo.snr_db(out, ref) >= 70.0, f"recording {i} of {Ts[i]} samples"
    again = plan.chain_batch(pcms)
    for a, b in zip(outs, again):
        assert np.array_equal(a, b)
    plan.close()


def test_ws_chain_unaligned_and_odd_channels(prototypes):
    """Channel counts that are not multiples of four take the scalar staging path; zero-length recordings emit nothing."""
    M, m, r = 512, 2, 2
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    rng = np.random.default_rng(11)
    for C in (1, 3, 6, 13):
        plan = btk_b200.Plan(M, m, r, C, h, g)
        W = _weights(rng, geo, C, M) if C > 1 else bo.ds_weights(np.zeros(1), FS, M)
        plan.set_weights(W)
        pcms = [wl.noise_recording(T, C, seed=T, sigma=300.0) for T in (0, 4000, 777)]
        outs = plan.chain_batch(pcms)
        for x, out in zip(pcms, outs):
            ref = bo.chain(x, h, g, geo, W)[2] if x.shape[0] else np.zeros(0, np.float32)
            assert out.shape == ref.shape
            if ref.size:
                assert bo.snr_db(out, ref) >= 70.0
        plan.close()


def test_raw_pointer_entry_points_validate_their_arrays(prototypes):
    """The batch calls hand (pointer, T) pairs to the C side: wrong shapes, dtypes or strides are refused before that."""
    M, m, r, C = 256, 4, 1, 4
    h, g = proto(prototypes, M, m, r)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    plan.set_ds_weights(FS, np.zeros(C))
    T = 1000
    good = wl.noise_recording(T, C, seed=1)
    out = np.empty(plan.chain_frames(T) * plan.D, np.float32)
    bad_inputs = [good[:, :3].copy(), good.reshape(-1), good.astype(np.float64), good[::2], good.T.copy().T]
    for bad in bad_inputs:
        for call in (plan.chain_batch_into, plan.mvdr_chain_batch_into, plan.chain_zelinski_batch_into):
            with pytest.raises(btk_b200.BtkError) as e:
                call([bad], [out])
            assert e.value.code == btk_b200._capi.EINVAL
    for bad_out in (out[:-1].copy(), out.astype(np.float64), np.empty(2 * out.size, np.float32)[::2]):
        with pytest.raises(btk_b200.BtkError) as e:
            plan.chain_batch_into([good], [bad_out])
        assert e.value.code == btk_b200._capi.EINVAL
    with pytest.raises(btk_b200.BtkError):
        plan.chain_batch_pcm_into([good.astype(np.int16)], btk_b200._capi.PCM_S16, [T + 1], [out])
    with pytest.raises(btk_b200.BtkError):
        plan.beamform(np.zeros((5, plan.B, C + 1), np.complex64))
    with pytest.raises(btk_b200.BtkError):
        plan.synthesis(np.zeros((5, plan.B + 1), np.complex64))
    plan.chain_batch_into([good], [out])          # the good arrays still pass
    plan.close()
