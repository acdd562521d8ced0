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


# ---------------------------------------------------------------------------------- ingest stream nodes
needs_ref = pytest.mark.skipif(not bo.CompiledReference.available(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
def test_ingest_restatements_pinned_to_reference_code():
    """The reference's own IterativeSampleFeature::next, Conversion24bit2Float::next and ChannelExtractionFeature::next
    (feature/feature.cc:868-896, 190-217, 3885-3900 -- extracted from the source at oracle build time, feature.cc itself
    needs libsndfile) against the restatements the device paths are compared with: bit-exact."""
    ref = bo.CompiledReference()
    rng = np.random.default_rng(0)
    for T, C, fs, bl, cfrom, cto in [(5000, 3, 1000, 16, 0, -1), (30016 * 2 + 5, 2, 1000, 16, 0, -1), (30016, 2, 1000, 16, 0, -1),
                                     (70000, 4, 1000, 32, 100, 40000), (100, 1, 8000, 128, 0, -1), (0, 2, 1000, 16, 0, -1)]:
        x = rng.integers(-32768, 32768, (T, C)).astype(np.float32)
        a, ta = bo.iterative_sample_blocks(x, fs, bl, cfrom, cto)
        b, tb = ref.iterative_sample(x, fs, bl, cfrom, cto)
        assert a.shape == b.shape and np.array_equal(a, b) and ta == tb
        assert a.shape[0] % (30 * fs // bl + 1) == 0                   # whole 30-s buffers, zero padded
    v = rng.integers(-(1 << 23), 1 << 23, 3000)
    v[:4] = [-(1 << 23), (1 << 23) - 1, -1, 0]
    raw = _pack_s24be(v)
    assert np.array_equal(ref.conversion24(raw, 100), bo.ingest_s24be(raw))
    d = rng.normal(size=(400, 5)).astype(np.float32)
    assert np.array_equal(ref.channel_extraction(d, 2, 5, 40), d[:, 2])


def _write_wav(path, pcm_s16, fs):
    import wave
    with wave.open(str(path), "wb") as w:
        w.setnchannels(pcm_s16.shape[1]); w.setsampwidth(2); w.setframerate(fs)
        w.writeframes(np.ascontiguousarray(pcm_s16, "<i2").tobytes())


@pytest.mark.gpu
def test_iterative_sample_feature_nodes(tmp_path, prototypes):
    """IterativeSampleFeaturePtr over a 16-bit WAVE file: lock-step pulling equals the pinned restatement block for block;
    whole_stream() is the same data per channel; banks on these nodes run the fused chain and match the oracle."""
    M, m, r, C, fs = 256, 4, 1, 4, 16000
    D = M >> r
    h, g = proto(prototypes, M, m, r)
    tau = wl.farfield_delays(wl.linear_array(C, 41.0), np.deg2rad(30), np.deg2rad(90))
    pcm16 = np.clip(np.round(wl.array_recording(20000, tau, seed=4, noise_sigma=400.0)), -32768, 32767).astype(np.int16)
    wav = tmp_path / "array.wav"
    _write_wav(wav, pcm16, fs)
    nodes = [btk_b200.IterativeSampleFeaturePtr(c, D, 0) for c in range(C)]
    for n in nodes:
        n.reset()
        n.read(str(wav))
    want, ttl = bo.iterative_sample_blocks(pcm16.astype(np.float32), fs, D)
    got = []
    try:
        while True:
            got.append(np.stack([np.array(n.next(), copy=True) for n in nodes]))
    except StopIteration:
        pass
    got = np.stack(got)
    assert got.shape == want.shape and np.array_equal(got, want) and nodes[0].samplesN() == ttl == 20000
    for c, n in enumerate(nodes):
        n.reset()
        assert np.array_equal(n.whole_stream(), want[:, c, :].reshape(-1))
    with pytest.raises(IOError):
        btk_b200.IterativeSampleFeaturePtr(0, D, 0).read(str(tmp_path / "missing.wav"))
    # the chain on these nodes (cut to the first second of the 30-s buffer to keep the oracle quick)
    bf = btk_b200.SubbandDSPtr(M)
    srcs = [btk_b200.IterativeSampleFeaturePtr(c, D, 0) for c in range(C)]
    for s in srcs:
        s.reset()
        s.read(str(wav), cfrom=0, cto=-1)
        bf.setChannel(btk_b200.OverSampledDFTAnalysisBankPtr(s, h, M, m, r))
    bf.calcArrayManifoldVectors(float(fs), tau)
    syn = btk_b200.OverSampledDFTSynthesisBankPtr(bf, g, M, m, r)
    out = np.concatenate([np.array(b, copy=True) for b in syn])
    assert syn.fused()
    full = want.transpose(0, 2, 1).reshape(-1, C)                          # [T'][C], zero padded to the buffer
    ref_out = bo.chain(full[:24000], h, g, bo.BankGeometry(M, m, r, 0), bo.ds_weights(tau, fs, M))[2]
    assert out.size == full.shape[0] and bo.snr_db(out[:20000], ref_out[:20000]) >= 70.0


@pytest.mark.gpu
def test_conversion24_and_channel_extraction_nodes():
    rng = np.random.default_rng(2)
    v = rng.integers(-(1 << 23), 1 << 23, 64 * 50)
    raw = _pack_s24be(v).reshape(-1)

    class Bytes(btk_b200.streams.FeatureStream):
        def __init__(self, data, block):
            super().__init__(block, "bytes")
            self._d, self._cur = data, 0

        def reset(self):
            super().reset()
            self._cur = 0

        def next(self, frameX=-5):
            if self._cur + self._size > self._d.size:
                raise btk_b200.streams.jiterator_error("end of samples!")
            b = self._d[self._cur:self._cur + self._size]
            self._cur += self._size
            self._frameX += 1
            return b

    conv = btk_b200.Conversion24bit2FloatPtr(Bytes(raw, 3 * 64))            # one Mark-III frame (64 ch x 3 bytes) per block
    assert conv.size() == 64
    frames = np.stack([np.array(b, copy=True) for b in conv])
    assert np.array_equal(frames.reshape(-1), v.astype(np.float32))
    conv.reset()
    ch = btk_b200.ChannelExtractionFeaturePtr(conv, 5, 64)
    assert ch.size() == 1
    assert np.array_equal(np.concatenate([np.array(b, copy=True) for b in ch]), v.reshape(-1, 64)[:, 5].astype(np.float32))


@pytest.mark.gpu
def test_cpp_iterative_sample_feature_chain(tmp_path, prototypes):
    """The C++ nodes: IterativeSampleFeature (WAVE reader + device widening) per channel -> banks -> SubbandDS -> synthesis."""
    import os
    import struct
    import subprocess
    from conftest import ROOT
    exe = tmp_path / "test_streams"
    libdir = os.path.join(ROOT, "distantspeechrecognition-mirror_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "host", "test_streams.cc"), "-o", str(exe),
                    f"-L{libdir}", "-lbtkb200", f"-Wl,-rpath,{libdir}"], check=True)
    M, m, r, C, fs, T = 256, 4, 1, 3, 16000, 9000
    D = M >> r
    h, g = proto(prototypes, M, m, r)
    mp = wl.linear_array(C, 41.0)
    tau = wl.farfield_delays(mp, np.deg2rad(30), np.deg2rad(90))
    pcm16 = np.clip(np.round(wl.array_recording(T, tau, seed=8, noise_sigma=400.0)), -32768, 32767).astype(np.int16)
    wav = tmp_path / "a.wav"
    _write_wav(wav, pcm16, fs)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("8i", M, m, r, 0, C, T, 5, 0))
        for a in (h, g, tau, mp):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
        f.write(struct.pack("d", 0.0))
        f.write(np.ascontiguousarray(pcm16, np.float32).tobytes())
    res = subprocess.run([str(exe), "chain", fin, fout, str(wav)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    raw = open(fout, "rb").read()
    n_out, fused, n_y, n_lock = struct.unpack("4i", raw[:16])
    out = np.frombuffer(raw, np.float32, n_out, 16)
    lock = np.frombuffer(raw, np.float32, n_lock, 16 + 4 * n_out + 8 * n_y)
    want, _ = bo.iterative_sample_blocks(pcm16.astype(np.float32), fs, D)
    assert fused == 1 and n_out == want.shape[0] * D
    assert np.array_equal(lock, want[:2, 1, :].reshape(-1))                 # lock-step blocks of channel 1
    full = want.transpose(0, 2, 1).reshape(-1, C)
    ref_out = bo.chain(full[:12000], h, g, bo.BankGeometry(M, m, r, 0), bo.ds_weights(tau, fs, M))[2]
    assert bo.snr_db(out[:9000], ref_out[:9000]) >= 70.0
