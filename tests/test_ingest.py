"""Raw-PCM ingest: int16 and packed big-endian 24-bit interleaved samples -> float32 on the device,
then the unchanged fused chain.  The conversion is integer -> float and therefore bit-exact against the oracle
(feature/feature.cc:190-217, 273, 868-896)."""
import numpy as np
import pytest

import btk_b200
import btk_oracle as bo
from conftest import proto

wl = btk_b200.workloads
cap = btk_b200._capi


def _pack_s24be(v):
    v = np.asarray(v, dtype=np.int64) & 0xFFFFFF
    return np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=-1).astype(np.uint8)


def test_oracle_s24be_known_answers():
    raw = np.array([[0x00, 0x00, 0x01], [0x7F, 0xFF, 0xFF], [0x80, 0x00, 0x00], [0xFF, 0xFF, 0xFF], [0x12, 0x34, 0x56]],
                   np.uint8)
    assert bo.ingest_s24be(raw).tolist() == [1.0, 8388607.0, -8388608.0, -1.0, float(0x123456)]
    v = np.random.default_rng(1).integers(-(1 << 23), 1 << 23, 1000)
    assert np.array_equal(bo.ingest_s24be(_pack_s24be(v)), v.astype(np.float32))
    assert bo.ingest_s16(np.array([-32768, -1, 0, 32767], np.int16)).tolist() == [-32768.0, -1.0, 0.0, 32767.0]


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 1023, 64 * 3001])
def test_convert_pcm_bit_exact(n, prototypes):
    h, g = proto(prototypes, 256, 4, 1)
    plan = btk_b200.Plan(256, 4, 1, 1, h, g)
    rng = np.random.default_rng(n)
    s16 = rng.integers(-32768, 32768, n).astype(np.int16)
    assert np.array_equal(plan.convert_pcm(s16, cap.PCM_S16), bo.ingest_s16(s16))
    v24 = rng.integers(-(1 << 23), 1 << 23, n)
    if n:
        v24[0] = -(1 << 23); v24[-1] = (1 << 23) - 1
    raw = _pack_s24be(v24)
    assert np.array_equal(plan.convert_pcm(raw, cap.PCM_S24BE), bo.ingest_s24be(raw))
    plan.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["s16", "s24be"])
def test_chain_from_raw_pcm_matches_float_path_and_oracle(fmt, prototypes):
    M, m, r, C = 256, 4, 1, 8
    h, g = proto(prototypes, M, m, r)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    tau = wl.farfield_delays(wl.circular_array(C), 1.0, 1.4)
    plan.set_ds_weights(16000.0, tau)
    rng = np.random.default_rng(5)
    Ts = [1, 777, 16000, 4099, 128 * 40]           # ragged batch, odd sizes (alignment padding between recordings)
    raws, floats = [], []
    for T in Ts:
        if fmt == "s16":
            x = np.round(wl.noise_recording(T, C, seed=int(rng.integers(1 << 30)), sigma=3000.0)).clip(-32768, 32767)
            raws.append(x.astype(np.int16)); floats.append(bo.ingest_s16(raws[-1]))
        else:
            v = rng.integers(-(1 << 23), 1 << 23, (T, C))
            raws.append(_pack_s24be(v)); floats.append(bo.ingest_s24be(raws[-1]))
    code = cap.PCM_S16 if fmt == "s16" else cap.PCM_S24BE
    outs = plan.chain_batch_pcm(raws, code)
    ref_dev = plan.chain_batch(floats)
    geo = bo.BankGeometry(M, m, r, 0)
    W = bo.ds_weights(tau, 16000.0, M)
    for o, rd, x in zip(outs, ref_dev, floats):
        assert np.array_equal(o, rd)               # same kernel, same float input: identical
        assert bo.snr_db(o, bo.chain(x, h, g, geo, W)[2]) >= 70.0
    plan.close()
