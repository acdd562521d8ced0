"""Drop-in boundary, second half (VERDICT r1 "What's missing" 2, 3, 7):

* the host nodes compiled INSIDE the reference tree (BTKB200_WITH_BTK: they derive from the reference's own
  stream/stream.h, common/refcount.h, common/jexception.h and beamformer/beamformer.h classes) and chained with the
  reference's own nodes -- oracle/_ref/mixed_chain, built by oracle/Makefile from tests/host/test_mixed_chain.cc;
* SnapShotArray / SpectralMatrixArray (beamformer/spectralinfoarray.h:6-67) and getSnapShotArray();
* calcArrayManifoldVectors2 / N (beamformer.cc:1100-1121, 603-735);
* G1: calcDelaysPolar2 (src/superdirectiveBeamformer.cc:118-137) and calcAllDelays (beamformer.cc:1214-1231).

CPU tier: what needs no device (oracle restatements against the compiled reference, the host-only C-ABI helpers, the
mixed binary's type checks).  GPU tier: the mixed chains and the device-backed classes.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

import btk_b200
import btk_oracle as bo
from conftest import ROOT, proto

wl = btk_b200.workloads
FS = 16000.0
MIXED = os.path.join(ROOT, "oracle", "_ref", "mixed_chain")
needs_ref = pytest.mark.skipif(not bo.CompiledReference.available(), reason="oracle/_ref not built (needs /root/reference)")
needs_mixed = pytest.mark.skipif(not os.path.exists(MIXED), reason="oracle/_ref/mixed_chain not built (needs /root/reference)")


@pytest.fixture(scope="module")
def ref():
    return bo.CompiledReference()


def geometry(C, jitter_seed=1):
    mp = np.zeros((C, 3))
    mp[:, 0] = np.arange(C) * 41.0
    mp[:, 1] = np.random.default_rng(jitter_seed).normal(0.0, 5.0, C)
    return mp


def interferers(mp, NC):
    return np.stack([bo.farfield_delays(mp, np.deg2rad(100 + 25 * n), np.deg2rad(80)) for n in range(NC - 1)])


# ---------------------------------------------------------------------------------------------- CPU tier
@needs_ref
def test_delay_helpers_match_compiled_reference(ref):
    """G1.  The C-ABI helpers (host arithmetic, no device) are bit-exact against the reference's own functions: calcAllDelays
    as compiled from beamformer.cc, calcDelaysPolar2 as extracted from the driver source at build time."""
    for C in (3, 8, 64):
        mp = geometry(C, C)
        mp[:, 2] = np.linspace(-20, 30, C)
        for az, el in ((0.7, 1.3), (2.1, np.float32(np.pi / 2)), (-1.0, 0.3)):
            assert np.array_equal(btk_b200.calc_delays_polar(az, el, mp), ref.delays_polar2(az, el, mp))
            # the numpy restatement: float32 sin/cos may differ from glibc's sinf/cosf in the last place
            d_ref = ref.delays_polar2(az, el, mp)
            assert np.abs(bo.delays_polar2(az, el, mp) - d_ref).max() <= 2.5e-7 * np.abs(d_ref).max()
        d = btk_b200.calc_all_delays(100.0, -50.0, 7.0, mp)            # the source position is ignored (beamformer.cc:1219-1223)
        assert np.array_equal(d, ref.all_delays(mp, 1.0, 2.0, 3.0)) and np.array_equal(d, btk_b200.calc_all_delays(0, 0, 0, mp))
        assert np.allclose(bo.all_delays(mp), d, rtol=0, atol=1e-18) and d[C // 2] == 0.0
    with pytest.raises(btk_b200.BtkError):
        btk_b200.calc_delays_polar(0.0, 0.0, np.zeros((4, 2)))


@needs_ref
@pytest.mark.parametrize("C,M,NC", [(4, 256, 2), (8, 512, 2), (6, 256, 3), (8, 64, 4)])
def test_null_weights_restatement_pinned(ref, C, M, NC):
    """bo.null_weights against SubbandDS::calcArrayManifoldVectors2 / N of the compiled reference, the bin-M/2 loop included.
    NC = 2 uses the closed-form 2 x 2 inverse on both sides: 1e-12.  For NC > 2 the reference inverts C^H C with its
    single-precision SVD, so it carries cond(C^H C) * float eps of error (measured: 0.6 absolute at bin 1, cond 2e6); the
    restatement (double) is gated per bin at 1e-6 * cond."""
    mp = geometry(C)
    dT = bo.farfield_delays(mp, np.deg2rad(30), np.deg2rad(90))
    dJ = interferers(mp, NC)
    w, ta = bo.null_weights(dT, dJ, FS, M)
    wr, tar = ref.null_weights(dT, dJ, FS, M)
    assert bo.rel_l2(ta, tar) <= 1e-15 and bo.rel_l2(ta, bo.ds_weights(dT, FS, M)) <= 1e-15   # the manifold stays delay-and-sum
    assert np.abs(w[0] - 1.0 / C).max() == 0.0
    if NC == 2:
        assert np.abs(w - wr).max() <= 1e-12
    else:
        for s in range(1, M // 2):
            pW = np.exp(1j * (-2 * np.pi * s * FS * dJ / M))
            Cm = np.column_stack([ta[s] * C] + [pW[n] for n in range(NC - 1)])
            cond = np.linalg.cond(Cm.conj().T @ Cm)
            assert np.abs(w[s] - wr[s]).max() <= 1e-6 * cond + 1e-6, (s, cond)
    # unit gain towards the target, nulls towards the interferers (what the constraint asks for), away from the
    # ill-conditioned lowest bins
    for s in range(M // 8, M // 2):
        v = np.exp(1j * (-2 * np.pi * s * FS * dT / M))
        assert abs(np.vdot(w[s], v) - 1.0) <= 1e-8
        for n in range(NC - 1):
            assert abs(np.vdot(w[s], np.exp(1j * (-2 * np.pi * s * FS * dJ[n] / M)))) <= 1e-8


@needs_ref
def test_end_of_stream_code_is_jiterator(ref):
    """common/jexception.h:41-57: JITERATOR is 8 (JPYTHON is 9); include/jexception.i:63-69 maps it to StopIteration."""
    assert ref.error_probe(3) == 8


@needs_mixed
def test_bound_nodes_use_the_reference_types():
    """BTKB200_WITH_BTK build: B200 nodes are held by refcountable_ptr, throw the reference's j_error family, a B200 bank
    is accepted by the reference's SubbandDS::setChannel and a B200 SubbandDS by ZelinskiPostFilter::setBeamformer."""
    r = subprocess.run([MIXED, "errors"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "errors: 0 failure(s)" in r.stdout


def test_snapshot_array_shadow():
    """SnapShotArrayPtr: newSample / update transpose (beamformer.cc:76-90) and newSnapShot's mirror (:99-113)."""
    sa = btk_b200.SnapShotArrayPtr(8, 3)
    X = np.arange(24).reshape(3, 8) + 1j * np.arange(24).reshape(3, 8)[::-1]
    for c in range(3):
        sa.newSample(X[c], c)
    sa.update()
    assert all(np.array_equal(sa.getSnapShot(s), X[:, s]) for s in range(8))
    sa.zero()
    assert not sa.getSnapShot(2).any()
    sa.newSnapShot(np.array([1 + 2j, 3 - 1j, 0.5j]), 1)
    assert np.array_equal(sa.getSnapShot(3), np.conj(sa.getSnapShot(1)))        # fftLen2 - fbinX, as the reference indexes it
    with pytest.raises(btk_b200.streams.jdimension_error):
        sa.newSample(np.zeros(7), 0)


# ---------------------------------------------------------------------------------------------- GPU tier
def write_input(path, M, m, r, dct, C, T, h, g, tau, pcm):
    with open(path, "wb") as f:
        f.write(struct.pack("8i", M, m, r, dct, C, T, 0, 0))
        for a in (h, g, tau):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
        f.write(np.ascontiguousarray(pcm, np.float32).tobytes())


@pytest.mark.gpu
@needs_mixed
@pytest.mark.parametrize("cfg", [(256, 4, 1, 0, 4, 9000), (512, 2, 2, 0, 3, 7000)])
def test_mixed_chains_match_all_reference_chain(cfg, tmp_path, prototypes):
    """Reference nodes and B200 nodes in one chain (oracle/_ref/mixed_chain run): the reference's own ZelinskiPostFilter
    between a B200 SubbandDS and a B200 synthesis bank, the reference's synthesis bank on a B200 beamformer, the
    reference's analysis banks under a B200 beamformer -- each against the all-reference chain, >= 70 dB."""
    M, m, r, dct, C, T = cfg
    h, g = proto(prototypes, M, m, r)
    mp = wl.linear_array(C, 41.0)
    tau = wl.farfield_delays(mp, np.deg2rad(30), np.deg2rad(90))
    pcm = wl.array_recording(T, tau, seed=77 + M, noise_sigma=700.0)
    fin = str(tmp_path / "in.bin")
    write_input(fin, M, m, r, dct, C, T, h, g, tau, pcm)
    res = subprocess.run([MIXED, "run", fin], capture_output=True, text=True)
    lines = [ln.split() for ln in res.stdout.splitlines() if ln[:3] in ("zel", "ds_")]
    assert res.returncode == 0, res.stdout + res.stderr
    got = {ln[0]: (float(ln[1]), int(ln[2]), ln[3]) for ln in lines}
    assert set(got) == {"zel_B_R_B", "ds_B_B_R", "ds_R_B_B", "ds_R_B_R", "ds_B_B_B"}
    nblk = -(-T // (M >> r))
    for name, (snr, frames, fused) in got.items():
        assert snr >= 70.0 and frames == nblk, (name, snr, frames)
    assert got["ds_B_B_B"][2] == "fused=1" and got["zel_B_R_B"][2] == "fused=0"


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("C,M,m,r,NC", [(4, 256, 4, 1, 2), (6, 512, 2, 2, 3)])
def test_null_steering_weights_on_the_device_path(ref, prototypes, C, M, m, r, NC):
    """btkb200_set_null_weights: the installed weights match the compiled reference (NC = 2: 1e-12), the array manifold
    stays delay-and-sum, and the fused chain with them matches the oracle chain with the reference's weights."""
    mp = geometry(C)
    dT = bo.farfield_delays(mp, np.deg2rad(30), np.deg2rad(90))
    dJ = interferers(mp, NC)
    h, g = proto(prototypes, M, m, r)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    plan.set_null_weights(FS, dT, dJ)
    w = plan.get_weights()
    wr, tar = ref.null_weights(dT, dJ, FS, M)
    assert bo.rel_l2(plan.get_array_manifold(), tar) <= 1e-15
    # NC > 2: two double-precision inverses of an ill-conditioned C^H C (cond 2e6 at bin 1) agree to cond * 1e-16
    assert np.abs(w - bo.null_weights(dT, dJ, FS, M)[0]).max() <= (1e-9 if NC == 2 else 1e-6)
    if NC == 2:
        assert np.abs(w - wr).max() <= 1e-12
    pcm = wl.array_recording(6000, dT, seed=11, noise_sigma=500.0)
    out = plan.chain(pcm)
    refout = bo.chain(pcm, h, g, bo.BankGeometry(M, m, r, 0), w)[2]
    assert out.shape == refout.shape and bo.snr_db(out, refout) >= 70.0
    with pytest.raises(btk_b200.BtkError):
        plan.set_null_weights(FS, dT, np.zeros((C, C)))                # NC = C + 1 > C
    plan.close()
    # the stream node
    bf = btk_b200.SubbandDSPtr(M)
    for c in range(C):
        src = btk_b200.SampleFeaturePtr(pcm[:, c], M >> r, M >> r, True)
        bf.setChannel(btk_b200.OverSampledDFTAnalysisBankPtr(src, h, M, m, r))
    bf.calcArrayManifoldVectorsN(FS, dT, dJ, NC)
    assert np.abs(bf.getWeights(5) - w[5]).max() <= 1e-15
    bf.next()
    sa = bf.getSnapShotArray()
    assert sa.nChan() == C and np.array_equal(sa.getSnapShot(7), bf.snapShotArray_f(7))
    assert np.array_equal(sa.getSnapShot(M - 7), np.conj(sa.getSnapShot(7)))


@pytest.mark.gpu
@needs_ref
def test_spectral_matrix_array_shadow_matches_compiled_reference(ref, prototypes):
    """SpectralMatrixArrayPtr (device-evaluated, lazily) against the reference's own SpectralMatrixArray fed by its own
    analysis banks: all M bins, an intermediate read in the middle of the stream."""
    M, m, r, C, T = 256, 4, 1, 5, 6000
    h, _ = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    tau = wl.farfield_delays(wl.linear_array(C, 41.0), np.deg2rad(40), np.deg2rad(90))
    pcm = wl.array_recording(T, tau, seed=3, noise_sigma=900.0)
    Rref = ref.spectral_matrix(pcm, h, geo, 0.95)                         # [M][C][C]
    X = np.stack([bo.analysis(pcm[:, c], h, geo) for c in range(C)], 1)   # [F][C][M]
    sma = btk_b200.SpectralMatrixArrayPtr(M, C, 0.95)
    sma.zero()
    for f in range(X.shape[0]):
        for c in range(C):
            sma.newSample(X[f, c], c)
        sma.update()
        if f == 17:
            mid = sma.getSpecMatrix(9).copy()                             # folds 18 frames, the recursion continues
            assert bo.rel_l2(mid, bo.spectral_matrix_cpp(X[:18], 0.95)[9]) <= 1e-4
    got = np.stack([sma.getSpecMatrix(s) for s in range(M)])
    assert got.shape == Rref.shape and bo.rel_l2(got, Rref) <= 1e-4
    assert max(bo.rel_l2(got[s], Rref[s]) for s in (0, 1, M // 2, M // 2 + 1, M - 1)) <= 1e-4


@pytest.mark.gpu
def test_cpp_snapshot_and_spectral_matrix_arrays(tmp_path, prototypes):
    """The C++ classes (host/btk_streams.h): SubbandDS::calcArrayManifoldVectors2, getSnapShotArray(), getBeamformerWeightObject(),
    SpectralMatrixArray fed frame by frame with an intermediate read -- tests/host/test_streams.cc mode 4."""
    from test_host_streams import exe as _exe  # noqa: F401  (fixture function reused below)
    out = tmp_path / "test_streams"
    libdir = os.path.join(ROOT, "distantspeechrecognition-mirror_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "host", "test_streams.cc"), "-o", str(out),
                    f"-L{libdir}", "-lbtkb200", f"-Wl,-rpath,{libdir}"], check=True)
    M, m, r, dct, C, T = 256, 4, 1, 0, 4, 5000
    h, g = proto(prototypes, M, m, r)
    mp = wl.linear_array(C, 41.0)
    tau = wl.farfield_delays(mp, np.deg2rad(30), np.deg2rad(90))
    pcm = wl.array_recording(T, tau, seed=21, noise_sigma=800.0)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("8i", M, m, r, dct, C, T, 4, 0))
        for a in (h, g, tau, mp):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
        f.write(struct.pack("d", 0.0))
        f.write(np.ascontiguousarray(pcm, np.float32).tobytes())
    res = subprocess.run([str(out), "chain", fin, fout], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    raw = open(fout, "rb").read()
    n, ny, nr, nw = struct.unpack("4i", raw[:16])
    off = 16
    Y = np.frombuffer(raw, np.float64, ny, off).view(np.complex128).reshape(-1, M); off += 8 * ny
    R = np.frombuffer(raw, np.float64, nr, off).view(np.complex128).reshape(M, C, C); off += 8 * nr
    W = np.frombuffer(raw, np.float64, nw, off).view(np.complex128).reshape(M // 2 + 1, C)
    geo = bo.BankGeometry(M, m, r, dct)
    w_ref, _ = bo.null_weights(tau, tau[::-1], FS, M)
    assert np.abs(W - w_ref).max() <= 1e-9
    X = np.stack([bo.analysis(pcm[:, c], h, geo) for c in range(C)], 1)   # [F][C][M]
    assert n == X.shape[0]
    assert bo.rel_l2(Y, bo.beamform(X, w_ref)[:4]) <= 1e-4
    assert bo.rel_l2(R, bo.spectral_matrix_cpp(X, 0.95)) <= 1e-4
