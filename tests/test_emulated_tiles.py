"""CPU tier: run the library's own tile programs (csrc/chain_tile.cuh, staged_tiles.cuh -- the code the
sm_100a kernels execute) sequentially on the host through tests/emu/emu_chain.cc and compare with the
oracle.  Checks index arithmetic / framing for every transform size without a GPU.  The emulator is a test
harness only; the product never links it."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import btk_b200
import btk_oracle as bo
from conftest import ROOT, proto

wl = btk_b200.workloads


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu") / "libbtk_emu.so"
    src = os.path.join(ROOT, "tests", "emu", "emu_chain.cc")
    inc = os.path.join(ROOT, "distantspeechrecognition-mirror_b200", "csrc")
    # BTK_EMU_FLAGS: extra -D switches, to run the same tests on a tuning variant of the tile programs before GPU time is spent
    extra = os.environ.get("BTK_EMU_FLAGS", "").split()
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", f"-I{inc}"] + extra + ["-x", "c++", src, "-o", str(out)],
                   check=True)
    return ctypes.CDLL(str(out))


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


CASES = [  # M, m, r, dct, C, T, chunk
    (256, 4, 1, 0, 1, 2000, 64),
    (256, 4, 1, 0, 8, 3000, 20),
    (256, 4, 1, 2, 3, 1500, 100),
    (512, 2, 2, 0, 4, 2200, 33),
    (512, 2, 3, 1, 2, 1700, 50),
    (128, 2, 1, 0, 5, 900, 25),
    (64, 2, 1, 0, 6, 500, 25),
    (1024, 2, 2, 0, 2, 4000, 32),
    (512, 2, 0, 0, 3, 5000, 24),
]


@pytest.mark.parametrize("fast", [3, 1, 0])   # bit 0: compile-time m, bit 1: two frame pairs per warp
@pytest.mark.parametrize("case", CASES)
def test_emulated_chain_matches_oracle(case, fast, emu, prototypes):
    M, m, r, dct, C, T, chunk = case
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, dct)
    pcm = wl.noise_recording(T, C, seed=7)
    mp = wl.circular_array(C) if C > 1 else np.zeros((1, 3))
    W = bo.ds_weights(wl.farfield_delays(mp, 1.0, 1.4), 16000.0, M)
    _, _, ref = bo.chain(pcm, h, g, geo, W)
    out = np.zeros(geo.nblk(T) * geo.D, np.float32)
    Ts, z = np.array([T], np.int64), np.array([0], np.int64)
    Wc = np.ascontiguousarray(W, dtype=np.complex128)
    n = emu.emu_chain(M, m, r, dct, C, 1, vp(Ts), vp(pcm), vp(z), vp(out), vp(z), vp(h), vp(g), vp(Wc), 1, chunk, fast)
    assert n > 0
    assert bo.snr_db(out, ref) > 100.0   # north_star gate is 70 dB


@pytest.mark.parametrize("fast", [3, 1, 0, 7, 5])   # bit 2: the transform warps keep the synthesis side (no overlap-add warps)
@pytest.mark.parametrize("case", CASES)
def test_emulated_ws_chain_matches_oracle(case, fast, emu, prototypes):
    """The warp-specialised tile program (csrc/chain_ws.cuh): producer fill + stage rotation + compute side + the overlap-add
    warps' program (frames parked per lane, ring of chunks) where the shape has one (compile-time m, two lane groups)."""
    M, m, r, dct, C, T, chunk = case
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, dct)
    pcm = wl.noise_recording(T, C, seed=9)
    mp = wl.circular_array(C) if C > 1 else np.zeros((1, 3))
    W = bo.ds_weights(wl.farfield_delays(mp, 1.0, 1.4), 16000.0, M)
    _, _, ref = bo.chain(pcm, h, g, geo, W)
    out = np.zeros(geo.nblk(T) * geo.D, np.float32)
    Ts, z = np.array([T], np.int64), np.array([0], np.int64)
    Wc = np.ascontiguousarray(W, dtype=np.complex128)
    n = emu.emu_chain_ws(M, m, r, dct, C, 1, vp(Ts), vp(pcm), vp(z), vp(out), vp(z), vp(h), vp(g), vp(Wc), 1, chunk, fast)
    if n == -2:
        pytest.skip("no warp-specialised layout for this shape (stages exceed 227 KB): the library uses chain_tile")
    assert n > 0
    assert bo.snr_db(out, ref) > 100.0   # north_star gate is 70 dB


@pytest.mark.parametrize("fast", [1, 0])
@pytest.mark.parametrize("case", CASES[:6])
def test_emulated_staged_tiles_match_oracle(case, fast, emu, prototypes):
    M, m, r, dct, C, T, chunk = case
    chunk = (chunk + 15) // 16 * 16
    h, g = proto(prototypes, M, m, r)
    geo = bo.BankGeometry(M, m, r, dct)
    pcm = wl.noise_recording(T, C, seed=11)
    X = np.stack([bo.analysis(pcm[:, c], h, geo) for c in range(C)], axis=1)
    F, B = X.shape[0], geo.B
    snap = np.zeros((F, B, C, 2), np.float32)
    assert emu.emu_analysis(M, m, r, dct, C, T, vp(pcm), vp(snap), vp(h), chunk, fast) == F
    assert bo.rel_l2(snap.view(np.complex64)[..., 0], X[:, :, :B].transpose(0, 2, 1)) < 1e-5   # gate: 1e-4
    Y = X[:, 0, :]
    ref = bo.synthesis(Y, g, geo).reshape(-1)
    Yh = np.ascontiguousarray(Y[:, :B]).astype(np.complex64)
    out = np.zeros(geo.synthesis_frames(F) * geo.D, np.float32)
    emu.emu_synthesis(M, m, r, dct, F, vp(Yh), vp(out), vp(g), 1, chunk, fast)
    assert bo.snr_db(out, ref) > 100.0


@pytest.mark.parametrize("ncta", [1, 7, 148])
def test_persistent_schedule_covers_every_frame_once(emu, ncta):
    """chain_ws.cuh::WsSegs: the contiguous item shares of the CTAs of a launch, walked as per-recording segments, cover the
    output frames of the launched recordings exactly once, in order, and no CTA has more than one item above its share."""
    rng = np.random.default_rng(3)
    W = 32
    nblk = np.array([0, 1, 31, 32, 33, 7500, 0, 1250, 64, 5, 0], np.int32)
    nblk = np.concatenate([nblk, rng.integers(0, 400, 40).astype(np.int32)])
    n = len(nblk)
    for r0, r1 in [(0, n), (3, 9), (5, 6), (6, 7), (10, n)]:
        cover = [np.zeros(int(b), np.int32) for b in nblk]
        out = np.zeros(3 * 64, np.int32)
        items = []
        for cta in range(ncta):
            k = emu.emu_ws_segments(n, vp(nblk), W, r0, r1, cta, ncta, vp(out), 64)
            assert 0 <= k <= 64
            its = 0
            last = (-1, -1)
            for rec, j0, nj in out[:3 * k].reshape(-1, 3):
                assert r0 <= rec < r1 and nj > 0 and j0 % W == 0
                assert (rec, j0) > last          # ascending, one segment per recording
                last = (rec, j0)
                cover[rec][j0:j0 + nj] += 1
                its += -(-nj // W)
            items.append(its)
        for r in range(n):
            assert np.all(cover[r] == (1 if r0 <= r < r1 else 0))
        assert max(items) - min(items) <= 1


@pytest.mark.parametrize("shape", [([7500] * 16, 32, 7, 148), ([1250] * 64, 16, 7, 148), ([0, 3, 40, 5000, 0, 1, 77] * 9, 32, 7, 148),
                                   ([300] * 3, 16, 3, 148), ([100000], 32, 15, 148), ([5] * 1000, 32, 7, 148)])
def test_balanced_schedule(emu, shape):
    """host_tables.h::balance_ctas + WsSegs with explicit CTA boundaries: every frame once, every CTA within the iteration
    budget B, and B - 1 would not have been enough for the equal-item split either (cfg2: 26 iterations instead of 27)."""
    nblk, W, H, ncta = shape
    nblk = np.array(nblk, np.int32)
    n = len(nblk)
    q = max(W // 8, 1)
    begin = np.zeros(ncta + 1, np.int32)
    B = emu.emu_ws_balance(n, vp(nblk), q, W, H, ncta, vp(begin))
    assert B >= 1 and np.all(np.diff(begin) >= 0) and begin[-1] == sum(-(-int(b) // q) for b in nblk)
    cover = [np.zeros(int(b), np.int32) for b in nblk]
    out = np.zeros(3 * 2048, np.int32)
    worst = 0
    for cta in range(ncta):
        k = emu.emu_ws_segments_balanced(n, vp(nblk), q, vp(begin), cta, ncta, vp(out), 2048)
        assert 0 <= k <= 2048
        its = 0
        for rec, j0, nj in out[:3 * k].reshape(-1, 3):
            assert nj > 0
            cover[rec][j0:j0 + nj] += 1
            its += -(-(nj + H) // W)
        worst = max(worst, its)
    for r in range(n):
        assert np.all(cover[r] == 1)
    assert worst <= B
    total = int(nblk.sum())
    assert B >= -(-(total // ncta + H) // W) or total < ncta      # no budget below the average share can work
    if list(nblk) == [7500] * 16:
        assert B == 26
