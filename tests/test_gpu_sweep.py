"""GPU tier: shape sweep of the fused chain and the staged path against the oracle (BASELINE configs[4]: 4-64
channels x M = 128-1024 x batches), small lengths so that the float64 oracle finishes in seconds.  Covers every
transform size, every decimation factor the kernels are instantiated for, the compile-time-m (m = 2, 4) and the
generic (m = 1, 3) code paths, channel counts that are not multiples of the staging group, and ragged batches."""
import numpy as np
import pytest

import btk_b200
import btk_oracle as bo

pytestmark = pytest.mark.gpu
wl = btk_b200.workloads
FS = 16000.0

SHAPES = []
for M in (64, 128, 256, 512, 1024):
    for r in (0, 1, 2, 3):
        for m in (1, 2, 3, 4):
            R = 1 << r
            if M == 64 and R > 8:
                continue
            if M == 1024 and (m > 2 or r == 0):      # shared-memory budget (plan_create refuses the rest)
                continue
            if (M, r) in ((512, 0),) and m > 2:
                continue
            SHAPES.append((M, m, r))
# thin the sweep deterministically: every third shape, plus the BASELINE geometries
SHAPES = sorted(set(SHAPES[::3] + [(256, 4, 1), (512, 2, 2), (512, 2, 3), (128, 2, 1), (1024, 2, 1)]))


@pytest.mark.parametrize("shape", SHAPES)
def test_chain_shape_sweep(shape):
    M, m, r = shape
    rng = np.random.default_rng(M * 100 + m * 10 + r)
    C = int(rng.choice([1, 2, 3, 4, 5, 8, 12]))
    dct = int(rng.choice([0, 1, 2]))
    D = M >> r
    h, g = wl.kaiser_prototype(M, m, r)
    geo = bo.BankGeometry(M, m, r, dct)
    try:
        plan = btk_b200.Plan(M, m, r, C, h, g, dct=dct)
    except btk_b200.BtkError as e:
        assert e.code == btk_b200._capi.EUNSUPPORTED      # too large for shared memory: refused, never wrong
        pytest.skip(f"unsupported shape {shape}: {e.msg}")
    mp = wl.circular_array(C) if C > 1 else np.zeros((1, 3))
    tau = wl.farfield_delays(mp, 0.7, 1.3)
    W = bo.ds_weights(tau, FS, M) * np.exp(1j * rng.uniform(0, 2 * np.pi, (geo.B, C)))   # arbitrary complex weights
    plan.set_weights(W)
    Ts = [int(rng.integers(1, 6 * D)), int(rng.integers(20 * D, 40 * D)) + int(rng.integers(0, D))]
    pcms = [wl.noise_recording(T, C, seed=int(rng.integers(1 << 30)), sigma=700.0) for T in Ts]
    outs = plan.chain_batch(pcms)
    for pcm, out in zip(pcms, outs):
        X, Y, ref = bo.chain(pcm, h, g, geo, W)
        assert out.shape == ref.shape
        assert bo.snr_db(out, ref) >= 70.0
    # staged path on the longer recording
    snap = plan.analysis(pcms[1])
    X, Y, ref = bo.chain(pcms[1], h, g, geo, W)
    assert bo.rel_l2(snap, X[:, :, : geo.B].transpose(0, 2, 1)) <= 1e-4
    Yd = plan.beamform(snap)
    assert bo.rel_l2(Yd, Y[:, : geo.B]) <= 1e-4
    assert bo.snr_db(plan.synthesis(Yd), ref) >= 70.0
    plan.close()
