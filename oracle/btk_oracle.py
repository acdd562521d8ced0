"""oracle/btk_oracle.py -- CPU oracle, TEST INFRASTRUCTURE ONLY.

numpy (float64 / complex128) restatement of the reference's subband front end:
analysis bank -> SubbandDS / SubbandMVDR -> synthesis bank.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module; the product (``distantspeechrecognition-mirror_b200``)
never does and has no CPU fallback.

Pinning: the reference ships no golden vectors, so this
restatement is pinned against the reference ITSELF -- its own .cc files compiled
in place by ``oracle/Makefile`` into ``oracle/_ref/libbtk_ref.so`` -- in
``tests/test_oracle_golden.py`` (test_oracle_matches_compiled_reference) and through the committed fixtures under
``tests/golden/`` (made by ``tests/golden/make_golden.py`` from that library).  The same holds for the later rows:
Zelinski post-filter (``make_golden_zelinski.py``, reference postfilter.cc compiled in), SubbandGSC with fixed active
weights (``make_golden_gsc.py``), de Haan prototype design (``make_golden_design.py``, reference prototypeDesign.cc compiled
in).  ONE row is a restatement without running reference code behind it -- PARITY UNPINNED for it: the raw-PCM ingest
(``ingest_s16`` / ``ingest_s24be``), because feature/feature.cc needs libsndfile and does not build here; it is checked on
hand-computed known answers only.

Every function cites the reference lines it restates (paths relative to
/root/reference/btk).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

SSPEED = 343740.0  # mm/s, beamformer/beamformer.h:47


# --------------------------------------------------------------------------- geometry
@dataclass(frozen=True)
class BankGeometry:
    """Derived sizes of OverSampledDFTFilterBank (modulated/modulated.cc:101-104, 262-300)."""

    M: int
    m: int
    r: int
    dct: int = 0

    @property
    def R(self) -> int:
        return 1 << self.r

    @property
    def D(self) -> int:
        return self.M >> self.r

    @property
    def N(self) -> int:
        return self.M * self.m

    @property
    def B(self) -> int:
        return self.M // 2 + 1

    @property
    def pd_analysis(self) -> int:
        # modulated.cc:278-296 (synthesis == false)
        return {0: 2 * self.m - 1, 1: self.m * self.R - 1, 2: self.m * self.R - 1}.get(self.dct, 2 * self.m - 1)

    @property
    def pd_synthesis(self) -> int:
        # modulated.cc:278-296 (synthesis == true)
        return {0: 2 * self.m - 1, 1: self.m * self.R - 1, 2: self.m * self.R // 2}.get(self.dct, 2 * self.m - 1)

    @property
    def laN(self) -> int:
        return (self.m * self.R // 2 - 1) if self.dct == 2 else 0

    def nblk(self, T: int) -> int:
        # SampleFeature::next: ceil(T / D) blocks, last one zero padded (feature/feature.cc:627-641)
        return -(-T // self.D)

    def analysis_frames(self, T: int) -> int:
        # modulated.cc:461-516: nblk real + pd padded frames, first laN internal frames skipped
        return self.nblk(T) + self.pd_analysis - self.laN

    def synthesis_frames(self, F: int) -> int:
        # modulated.cc:626-642: pd frames are consumed by priming, then one output per input frame
        return max(F - self.pd_synthesis, 0)


# --------------------------------------------------------------------------- analysis
def analysis(x: np.ndarray, h: np.ndarray, geo: BankGeometry) -> np.ndarray:
    """OverSampledDFTAnalysisBank::next over a whole recording (modulated/modulated.cc:412-516).

    x: float samples [T] (converted to double exactly like _RealBuffer::nextSample(float),
    modulated.h:121-130).  Returns complex128 [F][M] with F = nblk + pd - laN.
      u_i[q] = sum_k h[q + M k] * x[(i+1) D - 1 - q - M k]     (:419-434, zero before t=0 and after the data)
      X_i[s] = sum_q u_i[q] exp(+j 2 pi s q / M)                (:439, unnormalised backward FFT)
    """
    x = np.asarray(x, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64)
    M, m, D, N = geo.M, geo.m, geo.D, geo.N
    assert h.shape == (N,)
    T = x.shape[0]
    nblk = geo.nblk(T)
    nint = nblk + geo.pd_analysis  # internal frames i = 0 .. nint-1
    xx = np.zeros(N + (nint + 1) * D, dtype=np.float64)
    xx[N : N + T] = x
    # window of frame i is xx[i*D + D : i*D + D + N] reversed (newest sample first)
    n0 = N + (np.arange(nint) + 1) * D - 1  # index of the newest sample of frame i
    idx = n0[:, None] - np.arange(N)[None, :]
    win = xx[idx] * h[None, :]
    u = win.reshape(nint, m, M).sum(axis=1)
    X = np.fft.ifft(u, axis=1) * M
    return X[geo.laN :]


# --------------------------------------------------------------------------- delay-and-sum
def ds_weights(delays: np.ndarray, fs: float, M: int) -> np.ndarray:
    """beamformerWeights::calcMainlobe, halfBandShift == false (beamformer/beamformer.cc:531-594).

    Returns wq [B][C] complex128: wq[0] = 1/C; wq[s] = exp(-j 2 pi s tau fs / M)/C for 1 <= s < M/2;
    wq[M/2] = exp(-j pi fs tau)/C.
    """
    tau = np.asarray(delays, dtype=np.float64)
    C = tau.shape[0]
    B = M // 2 + 1
    s = np.arange(B, dtype=np.float64)
    ph = -2.0 * np.pi * s[:, None] * tau[None, :] * fs / M
    ph[M // 2] = -np.pi * fs * tau
    w = np.exp(1j * ph) / C
    w[0] = 1.0 / C
    return w


def beamform(X: np.ndarray, W: np.ndarray) -> np.ndarray:
    """SubbandDS::next / SubbandMVDR::next (beamformer.cc:1137-1200, 2583-2635).

    X: [F][C][M] per-channel full spectra (only bins 0..M/2 are read, like the reference's
    snapshot loop), W: [B][C].  Y[s] = sum_c conj(W[s,c]) X_c[s] (gsl_blas_zdotc conjugates its
    first argument), Y[M-s] = conj(Y[s]).  Returns [F][M].
    """
    F, C, M = X.shape
    B = M // 2 + 1
    Yh = np.einsum("sc,fcs->fs", np.conj(W), X[:, :, :B])
    Y = np.empty((F, M), dtype=np.complex128)
    Y[:, :B] = Yh
    Y[:, B:] = np.conj(Yh[:, 1 : M // 2][:, ::-1])
    return Y


# --------------------------------------------------------------------------- MVDR
def diffuse_coherence(micpos_mm: np.ndarray, fs: float, M: int, sspeed: float = SSPEED) -> np.ndarray:
    """SubbandMVDR::setDiffuseNoiseModel (beamformer.cc:2486-2553): Gamma_mn = sinc(2 fs s d_mn/(M c)),
    GSL's normalised sinc == numpy's, unit diagonal.  Returns [B][C][C] complex128."""
    p = np.asarray(micpos_mm, dtype=np.float64)
    d = np.sqrt(((p[:, None, :3] - p[None, :, :3]) ** 2).sum(-1))
    B = M // 2 + 1
    s = np.arange(B, dtype=np.float64)
    G = np.sinc((2.0 * fs * s / (M * sspeed))[:, None, None] * d[None])
    i = np.arange(p.shape[0])
    G[:, i, i] = 1.0
    return G.astype(np.complex128)


def divide_nondiagonal(Rn: np.ndarray, mu: float) -> np.ndarray:
    """SubbandMVDR::divideAllNonDiagonalElements (beamformer.h:362-378); mu passes through float."""
    out = np.array(Rn, dtype=np.complex128, copy=True)
    C = out.shape[-1]
    off = ~np.eye(C, dtype=bool)
    out[:, off] = out[:, off] / (1.0 + float(np.float32(mu)))
    return out


def diagonal_load(Rn: np.ndarray, load: float) -> np.ndarray:
    """SubbandMVDR::setAllLevelsOfDiagonalLoading (beamformer.cc:2555-2568): the weight is stored in a
    float array (_diagonalWeights) before it is added, so it is rounded to float32 first."""
    out = np.array(Rn, dtype=np.complex128, copy=True)
    i = np.arange(out.shape[-1])
    out[:, i, i] += float(np.float32(load))
    return out


def mvdr_weights(Rn: np.ndarray, wq: np.ndarray) -> np.ndarray:
    """SubbandMVDR::calcMVDRWeights (beamformer.cc:2392-2446) with an exact inverse ("oracle B").

    w[0] = ones (NOT 1/C, :2410-2415); for s >= 1: t = Rinv^H d, lam = t^H d, w = t / (lam * C)
    with d = wq[s] (already divided by C).  Rn: [B][C][C], wq: [B][C]."""
    B, C = wq.shape
    w = np.ones((B, C), dtype=np.complex128)
    for s in range(1, B):
        Rinv = np.linalg.inv(Rn[s])
        t = Rinv.conj().T @ wq[s]
        lam = np.vdot(t, wq[s])  # conj(t) . d
        w[s] = t / (lam * C)
    return w


# --------------------------------------------------------------------------- covariance
def spectral_matrix_cpp(X: np.ndarray, mu: float = 0.95) -> np.ndarray:
    """SpectralMatrixArray::update over all frames (beamformer.cc:142-163): R <- mu R + (1-mu) x x^T,
    NO conjugate, all M bins, R starts at zero.  X: [F][C][M].  Returns [M][C][C]."""
    F, C, M = X.shape
    wts = (1.0 - mu) * mu ** np.arange(F - 1, -1, -1, dtype=np.float64)
    return np.einsum("f,fis,fjs->sij", wts, X, X)


def spectral_matrix_py(X: np.ndarray, ff: float = 0.99, nbins: int | None = None) -> np.ndarray:
    """SubbandBeamformerMVDR.updateSx (lib/subbandBeamforming.py:1170-1175): frame 0: S = x x^H, then
    S <- ff S + (1-ff) x x^H.  X: [F][C][M].  Returns [nbins][C][C] (default bins 0..M/2)."""
    F, C, M = X.shape
    nb = M // 2 + 1 if nbins is None else nbins
    wts = (1.0 - ff) * ff ** np.arange(F - 1, -1, -1, dtype=np.float64)
    wts[0] = ff ** (F - 1)
    Xb = X[:, :, :nb]
    return np.einsum("f,fis,fjs->sij", wts, Xb, np.conj(Xb))


# --------------------------------------------------------------------------- synthesis
def synthesis(Y: np.ndarray, g: np.ndarray, geo: BankGeometry, gain: int = 1) -> np.ndarray:
    """OverSampledDFTSynthesisBank::next over a whole stream (modulated/modulated.cc:595-664).

    Y: [F][M] full spectra.  Returns float32 [nout][D] with nout = F - pd.
      v_tau[q]  = Re sum_s Y_tau[s] exp(-j 2 pi s q / M)                       (:603-607)
      w_j[q]    = sum_k g[M-1-q + M k] v_{j+pd-R k}[q]   (v of frames < 0 is 0) (:646-651)
      out_j[D-1-d] = sum_{s<R} w_{j-(R-1-s)}[d + s D],  w_{j'<0} = 0            (:655-658, priming quirk:
                     the pd priming frames never produce a w, so the first R-1 outputs miss terms)
    The reference accumulates the R terms into a float vector (gsl_vector_float_set of float + double),
    so the sum is rounded to float32 after every term; that order is kept here.
    """
    Y = np.asarray(Y, dtype=np.complex128)
    g = np.asarray(g, dtype=np.float64)
    M, m, R, D = geo.M, geo.m, geo.R, geo.D
    F = Y.shape[0]
    pd = geo.pd_synthesis
    nout = geo.synthesis_frames(F)
    if nout == 0:
        return np.zeros((0, D), dtype=np.float32)
    v = np.real(np.fft.fft(Y, axis=1))
    vp = np.concatenate([np.zeros((m * R, M)), v], axis=0)  # vp[m*R + tau] = v_tau
    gp = g.reshape(m, M)[:, ::-1]  # gp[k, q] = g[M-1-q + M k]
    j = np.arange(nout)
    w = np.zeros((nout, M))
    for k in range(m):
        w += gp[k][None, :] * vp[m * R + j + pd - R * k]
    wp = np.concatenate([np.zeros((R, M)), w], axis=0)  # wp[R + j] = w_j
    out = np.zeros((nout, D), dtype=np.float32)
    for s in range(R):  # sampX order of modulated.cc:655-658
        term = wp[R + j - (R - 1 - s)][:, s * D : (s + 1) * D]
        out = (out.astype(np.float64) + term).astype(np.float32)
    out = out[:, ::-1]
    if gain > 0:
        out = (out * np.float32(gain)).astype(np.float32)
    return np.ascontiguousarray(out)


# --------------------------------------------------------------------------- geometry helpers (G1)
def farfield_delays(micpos_mm: np.ndarray, azimuth: float, elevation: float, sspeed: float = SSPEED) -> np.ndarray:
    """calcDelaysPolar2-style far-field delays (src/superdirectiveBeamformer.cc:118-137,
    lib/subbandBeamforming.py:227-246): tau_c = (c . p_c)/sspeed with
    c = -(sin(el) cos(az), sin(el) sin(az), cos(el)); geometry in mm."""
    p = np.asarray(micpos_mm, dtype=np.float64)
    c = -np.array([np.sin(elevation) * np.cos(azimuth), np.sin(elevation) * np.sin(azimuth), np.cos(elevation)])
    return (p[:, :3] @ c) / sspeed


# --------------------------------------------------------------------------- whole chain
def delays_polar2(azimuth: float, elevation: float, micpos_mm: np.ndarray) -> np.ndarray:
    """calcDelaysPolar2 of the shipped driver (src/superdirectiveBeamformer.cc:118-137): the direction cosines, the
    coordinates and the sum are float32 there, the quotient by SOUNDSPEED is taken in double and rounded to float32."""
    f = np.float32
    az, el = f(azimuth), f(elevation)
    c = (-(np.sin(el) * np.cos(az)).astype(f), -(np.sin(el) * np.sin(az)).astype(f), (-np.cos(el)).astype(f))
    mp = np.asarray(micpos_mm, dtype=np.float64).astype(f)
    s = (c[0] * mp[:, 0]).astype(f)
    s = (s + (c[1] * mp[:, 1]).astype(f)).astype(f)
    s = (s + (c[2] * mp[:, 2]).astype(f)).astype(f)
    return (s.astype(np.float64) / SSPEED).astype(f).astype(np.float64)


def all_delays(micpos_mm: np.ndarray) -> np.ndarray:
    """calcAllDelays (beamformer.cc:1214-1231): |position| / c minus the middle element's; its x, y, z arguments are
    never read (:1219-1223)."""
    mp = np.asarray(micpos_mm, dtype=np.float64)
    d = np.sqrt((mp[:, :3] ** 2).sum(1)) / SSPEED
    return d - d[mp.shape[0] // 2]


def _null_beamformer(wt: np.ndarray, pWj: np.ndarray) -> np.ndarray:
    """calcNullBeamformer (beamformer.cc:315-397): wt <- Cm (Cm^H Cm)^-1 e0, Cm = [wt | pWj...]; NC = 2 through
    putInverseMat22 (:202-242), larger NC through a double-precision inverse (the reference: float SVD pseudoinverse)."""
    Cm = np.column_stack([wt] + [pWj[n] for n in range(pWj.shape[0])])
    A = Cm.conj().T @ Cm
    if Cm.shape[1] == 2:
        m00, m01, m10, m11 = A[0, 0], A[0, 1], A[1, 0], A[1, 1]
        det = m00 * m11 - m01 * m10
        if abs(det) < 1.0e-7:
            m00, m11 = m00 + 0.01, m11 + 0.01
            det = m00 * m11 - m01 * m10
        v = np.array([m11 / det, -(m10 / det)])
    else:
        v = np.linalg.inv(A)[:, 0]
    return Cm @ v


def null_weights(delaysT: np.ndarray, delaysJ: np.ndarray, fs: float, M: int):
    """beamformerWeights::calcMainlobeN, halfBandShift = False (beamformer.cc:632-735): (quiescent weights, array manifold),
    both [B][C].  Statement for statement, the bin-M/2 loop (:722-734) included: element c is overwritten with the last
    interferer's steering value / C and the null beamformer re-run after EVERY channel, with the interferer vectors of
    bin M/2 - 1."""
    delaysT = np.asarray(delaysT, dtype=np.float64)
    dJ = np.atleast_2d(np.asarray(delaysJ, dtype=np.float64))
    C = delaysT.shape[0]
    ta = ds_weights(delaysT, fs, M)
    w = ta.copy()
    pWj = np.zeros((dJ.shape[0], C), dtype=np.complex128)
    for s in range(1, M // 2):
        vec = w[s] * C
        pWj = np.exp(1j * (-2.0 * np.pi * s * fs * dJ / M))
        w[s] = _null_beamformer(vec, pWj)
    vec = w[M // 2].copy()
    for c in range(C):
        vec[c] = vec[c] * C
        for n in range(dJ.shape[0]):
            vec[c] = np.exp(1j * (-np.pi * fs * dJ[n, c])) / C
        vec = _null_beamformer(vec, pWj)
    w[M // 2] = vec
    return w, ta


def chain(pcm: np.ndarray, h: np.ndarray, g: np.ndarray, geo: BankGeometry, W: np.ndarray, gain: int = 1):
    """analysis (per channel) -> beamform with weights W [B][C] -> synthesis.
    pcm: [T][C] float32.  Returns (X [F][C][M], Y [F][M], out float32 [nblk*D])."""
    pcm = np.asarray(pcm)
    C = pcm.shape[1]
    X = np.stack([analysis(pcm[:, c], h, geo) for c in range(C)], axis=1)
    Y = beamform(X, W)
    out = synthesis(Y, g, geo, gain)
    return X, Y, out.reshape(-1)


# --------------------------------------------------------------------------- prototype design (de Haan)
def _pinv_solve(K: np.ndarray, b: np.ndarray, tol: float) -> np.ndarray:
    """x = V diag(1/s_n or 0) U^T b with singular values below tol * s_0 dropped (prototypeDesign.cc:691-712, 885-901)."""
    U, s, Vt = np.linalg.svd(K)
    c = U.T @ b
    keep = (s / s[0]) >= tol
    c = np.where(keep, c / np.where(keep, s, 1.0), 0.0)
    return Vt.T @ c


def design_analysis_dehaan(M: int, m: int, r: int, wp_factor: float = 1.0, tau: int = -1, tol: float = 1e-7) -> np.ndarray:
    """AnalysisOversampledDFTDesign::design (prototypeDesign.cc:223-272, 640-712):  h = pinv(A + C) b with
    A(m,n) = sinc(wp (n-m)), b(m) = sinc(wp (tau - m)), wp = pi / (wpFactor M), tau = L/2 by default,
    C(m,n) = f/D (n == m) or f sin(pi (n-m)/D) / (pi D (n-m)), f = D-1 where D divides n-m, else -1."""
    L, D = M * m, M >> r
    wp = np.pi / (wp_factor * M)
    tau = L // 2 if tau < 0 else tau
    idx = np.arange(L)
    d = idx[None, :] - idx[:, None]                       # n - m
    with np.errstate(divide="ignore", invalid="ignore"):
        A = np.where(d == 0, 1.0, np.sin(wp * d) / (wp * d))
        t = tau - idx
        b = np.where(t == 0, 1.0, np.sin(wp * t) / (wp * t))
        f = np.where(d % D == 0, D - 1.0, -1.0)
        C = np.where(d == 0, f / D, f * np.sin(np.pi * d / D) / (np.pi * D * d))
    return _pinv_solve(A + C, b, tol)


def design_synthesis_dehaan(h: np.ndarray, M: int, m: int, r: int, v: float = 1.0, tau: int = -1,
                            tol: float = 1e-7) -> np.ndarray:
    """SynthesisOversampledDFTDesign::design (prototypeDesign.cc:836-901):  g = pinv(E + v P) f with
    E(m,n) = (M/D)^2 sum_{k=0..2m} h[kM-m] h[kM-n],  P(m,n) = M/D^2 f(m-n) sum_k h[k+n] h[k+m]  (f as above),
    f(m) = M/(pi D) h[2 tau - m]; out-of-range taps are skipped."""
    h = np.asarray(h, dtype=np.float64)
    L, D = M * m, M >> r
    tau = L // 2 if tau < 0 else tau
    idx = np.arange(L)
    E = np.zeros((L, L))
    for k in range(2 * m + 1):
        j = k * M - idx
        ok = (j >= 0) & (j <= L - 1)
        col = np.where(ok, h[np.clip(j, 0, L - 1)], 0.0)
        E += np.outer(col, col)
    E *= float((M // D) * (M // D))
    rr = np.correlate(h, h, mode="full")                   # rr[L-1+d] = sum_j h[j] h[j+d]
    d = idx[:, None] - idx[None, :]                        # m - n
    fac = np.where(d % D == 0, D - 1.0, -1.0)
    P = fac * rr[L - 1 + np.abs(d)] * (M / (float(D) * float(D)))
    j = 2 * tau - idx
    ok = (j >= 0) & (j <= L - 1)
    f = np.where(ok, h[np.clip(j, 0, L - 1)], 0.0) * (M / (np.pi * D))
    return _pinv_solve(E + v * P, f, tol)


def _design_matrices_analysis(M, m, r, wp_factor, tau):
    L, D = M * m, M >> r
    wp = np.pi / (wp_factor * M)
    tau = L // 2 if tau < 0 else tau
    idx = np.arange(L)
    d = idx[None, :] - idx[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        A = np.where(d == 0, 1.0, np.sin(wp * d) / (wp * d))
        t = tau - idx
        b = np.where(t == 0, 1.0, np.sin(wp * t) / (wp * t))
        f = np.where(d % D == 0, D - 1.0, -1.0)
        C = np.where(d == 0, f / D, f * np.sin(np.pi * d / D) / (np.pi * D * d))
    return A, b, C


def _design_matrix_P(h, M, m, r):
    L, D = M * m, M >> r
    idx = np.arange(L)
    rr = np.correlate(h, h, mode="full")
    d = idx[:, None] - idx[None, :]
    fac = np.where(d % D == 0, D - 1.0, -1.0)
    return fac * rr[L - 1 + np.abs(d)] * (M / (float(D) * float(D)))


def _svd_full(At: np.ndarray):
    """Singular values (decreasing) and right singular vectors V (all N of them, orthonormal) of At [rows x N] -- what
    PrototypeDesignBase::_svd hands back for either shape (prototypeDesign.cc:276-293)."""
    _, s, Vt = np.linalg.svd(At, full_matrices=True)
    N = At.shape[1]
    sv = np.zeros(N)
    sv[: s.size] = s
    return sv, Vt.T


def _constrained_design(Hc, c0, Q, A, b, tol):
    """The two solution paths of the Nyquist(M) designs (prototypeDesign.cc:361-470, 481-577, 579-609).
    Hc [L x k]: constraint columns (Hc^T x = c0), Q [L x L]: the matrix of the quadratic to minimise.
    cond([Hc^T; Q]) < 1/tol  -> "alternate solution 4" (_solveNonSingular): x = x_pt - B pinv_tol(B^T Q B) B^T Q x_pt over the
                                null space B of Hc^T;
    otherwise                -> "alternate solution 3" (_solveSingular): constraints K = [Hc Q] (K^T x = [c0; 0]), then
                                x = x_pt + B pinv'(B^T A B) B^T (b - A x_pt) over the numerical null space B of K^T, where
                                pinv' divides the components above the tolerance and LEAVES the others as they are (:556-558).
    Returns (x, path)."""
    L = Hc.shape[0]
    Kt = np.vstack([Hc.T, Q])
    sv, V = _svd_full(Kt)
    cond = sv[0] / sv[L - 1] if sv[L - 1] > 0 else np.inf
    if cond < 1.0 / tol:
        s, Vh = _svd_full(Hc.T)
        U = np.linalg.svd(Hc.T, full_matrices=False)[0]                 # [k x k]
        keep = (s / s[0]) >= tol
        coef = np.zeros(L)
        k = Hc.shape[1]
        coef[:k] = np.where(keep[:k], (U.T @ c0) / np.where(keep[:k], s[:k], 1.0), 0.0)
        x_pt = Vh @ coef
        B = Vh[:, L - int(np.sum((s / s[0]) < tol)):]
        Qt = B.T @ Q @ B
        Uq, sq, Vqt = np.linalg.svd(Qt)
        y = Uq.T @ (B.T @ (Q @ x_pt))
        y = np.where((sq / sq[0]) > tol, y / np.where(sq > 0, sq, 1.0), 0.0)
        return x_pt - B @ (Vqt.T @ y), 4
    # singular branch
    dp = np.concatenate([c0, np.zeros(L)])
    Uk, sk, Vkt = np.linalg.svd(Kt, full_matrices=False)                  # Kt: [(k + L) x L]
    keep = (sk / sk[0]) >= tol
    x_pt = Vkt.T @ np.where(keep, (Uk.T @ dp) / np.where(keep, sk, 1.0), 0.0)
    nnull = int(np.sum((sk / sk[0]) < tol))
    B = Vkt.T[:, L - nnull:]
    At_ = B.T @ A @ B
    Ua, sa, Vat = np.linalg.svd(At_)
    y = Ua.T @ (B.T @ (b - A @ x_pt))
    y = np.where((sa / sa[0]) > tol, y / np.where(sa > 0, sa, 1.0), y)
    return x_pt + B @ (Vat.T @ y), 3


def design_analysis_nyquist(M: int, m: int, r: int, wp_factor: float = 1.0, tau: int = -1, tol: float = 1e-7,
                            want_path: bool = False):
    """AnalysisNyquistMDesign::design (prototypeDesign.cc:955-1001): minimise the in-band aliasing h^T C h (or, when the
    constraints are numerically dependent, the passband error h^T A h - 2 h^T b inside their null space) subject to the
    Nyquist(M) constraint h[n M] = delta(n - m/2) / M  (F [L x m], F[n M, n] = 1; d[m/2] = 1/M, :982-989)."""
    L = M * m
    A, b, C = _design_matrices_analysis(M, m, r, wp_factor, tau)
    F = np.zeros((L, m))
    d = np.zeros(m)
    for n in range(m):
        F[n * M, n] = 1.0
    d[m // 2] = 1.0 / M
    x, path = _constrained_design(F, d, C, A, b, tol)
    return (x, path) if want_path else x


def design_synthesis_nyquist(h: np.ndarray, M: int, m: int, r: int, wp_factor: float = 1.0, tau: int = -1,
                             tol: float = 1e-7, want_path: bool = False):
    """SynthesisNyquistMDesign::design (prototypeDesign.cc:1003-1119): minimise the residual aliasing g^T P g subject to the
    2 m total-response constraints  H^T g = c0,  H[k, n] = h[n M - k] for max(0, 1 + (n - m) M) <= k <= min(n M, m M - 1)
    (:1073-1089), c0[m] = D / M; singular branch as in the analysis design with the passband matrices A, b."""
    h = np.asarray(h, dtype=np.float64)
    L, D = M * m, M >> r
    A, b, _ = _design_matrices_analysis(M, m, r, wp_factor, tau)
    P = _design_matrix_P(h, M, m, r)
    H = np.zeros((L, 2 * m))
    for n in range(2 * m):
        for k in range(max(0, 1 + (n - m) * M), min(n * M, m * M - 1) + 1):
            H[k, n] = h[n * M - k]
    c0 = np.zeros(2 * m)
    c0[m] = float(D) / M
    x, path = _constrained_design(H, c0, P, A, b, tol)
    return (x, path) if want_path else x


# --------------------------------------------------------------------------- SubbandGSC (fixed active weights)
def blocking_matrix(v: np.ndarray, NC: int = 1) -> np.ndarray:
    """_calcBlockingMatrix (beamformer/beamformer.cc:398-479), loop for loop: P = I - conj(v) v^T / ||v||^2 (zgeru), then
    Gram-Schmidt of its first C-NC columns with zdotc (first argument conjugated) and unit normalisation.  [C][C-NC]."""
    v = np.asarray(v, dtype=np.complex128)
    C = v.size
    bs = C - NC
    if bs <= 0:
        raise ValueError(f"The number of sensors {C} > the number of constraints {NC}")
    P = np.eye(C, dtype=np.complex128) - np.outer(np.conj(v), v) / (np.linalg.norm(v) ** 2)
    Bm = np.zeros((C, bs), dtype=np.complex128)
    for idim in range(bs):
        vec = P[:, idim].copy()
        for jdim in range(idim):
            rvec = Bm[:, jdim]
            vec = vec - np.vdot(rvec, vec) * rvec
        Bm[:, idim] = vec / np.linalg.norm(vec)
    return Bm


def gsc_weights(wq: np.ndarray, wa: np.ndarray, normalize: bool = False) -> np.ndarray:
    """The weights SubbandGSC::next applies (beamformer.cc:1296-1356 with calcOutputOfGSC :1251-1289):
    bin 0: wq alone; bins 1..M/2: w = wq - B wa, divided by ||w|| C when normalizeWeight is set.
    wq [B][C] quiescent (delay-and-sum) vectors, wa [B][C-1] active weights."""
    wq = np.asarray(wq, dtype=np.complex128)
    W = wq.copy()
    C = wq.shape[1]
    for s in range(1, wq.shape[0]):
        w = wq[s] - blocking_matrix(wq[s]) @ np.asarray(wa[s], dtype=np.complex128)
        if normalize:
            w = w / (np.linalg.norm(w) * C)
        W[s] = w
    return W


# --------------------------------------------------------------------------- Zelinski post-filter
TYPE_ZELINSKI1_REAL, TYPE_ZELINSKI1_ABS, NO_USE_POST_FILTER = 1, 2, 0      # postfilter/postfilter.h:66-72
SPECTRAL_FLOOR = 1.0e-4                                                     # postfilter/postfilter.cc:56


def zelinski_postfilter(X: np.ndarray, Y: np.ndarray, ta: np.ndarray, alpha: float = 0.6,
                        pf_type: int = TYPE_ZELINSKI1_ABS, min_frames: int = 0):
    """ZelinskiPostFilter::next over a stream (postfilter/postfilter.cc:428-500), restated with the reference's own
    per-pair state:  X [F][C][M] snapshots, Y [F][M] beamformer output, ta [B][C] array manifold
    (beamformerWeights::arrayManifold = the delay-and-sum weights, beamformer.cc:583, 992-997).
    Per frame n (the node's _frameX is n - 1 when the frame is processed):
      alpha_n = alpha if n - 1 > 0 else 0                                        (:466-469)
      y_i = conj(ta_i) x_i                                                        (TimeAlignment, :30-43)
      Phi_ij <- alpha_n Phi_ij + (1 - alpha_n) y_i conj(y_j), i < j  (or y_i conj(y_j) when alpha_n == 0)   (calcCSD, :8-21)
      Psi_i  <- alpha_n Psi_i  + (1 - alpha_n) |y_i|^2                            (:96-110)
      W = clamp(num / sum Psi * 2 / (C - 1), 1e-4, 1), num = max(Re sum Phi, 0) (REAL) or |sum Phi| (otherwise)  (:84-124)
      bins 0..M/2, mirrored; the signal is multiplied only when n - 1 >= min_frames and type != 0   (:474-479, 197-199)
    Returns (Ypf [F][M] complex128, W [F][B] float64)."""
    X = np.asarray(X, dtype=np.complex128)
    Y = np.asarray(Y, dtype=np.complex128)
    F, C, M = X.shape
    B = M // 2 + 1
    if C <= 1:
        raise ValueError(f"The number of channels {C} is <= 1")
    iu = np.triu_indices(C, 1)
    Phi = np.zeros((B, iu[0].size), dtype=np.complex128)
    Psi = np.zeros((B, C), dtype=np.float64)
    Ypf = Y.copy()
    Wall = np.zeros((F, B), dtype=np.float64)
    for n in range(F):
        a = alpha if (n - 1) > 0 else 0.0
        y = np.conj(ta[:B]) * X[n, :, :B].T                     # [B][C]
        cross = y[:, iu[0]] * np.conj(y[:, iu[1]])
        if a > 0.0:
            Phi = a * Phi + (1.0 - a) * cross
            Psi = a * Psi + (1.0 - a) * np.abs(y) ** 2
        else:
            Phi = cross
            Psi = np.abs(y) ** 2
        tot = Phi.sum(axis=1)
        # frames that only update the densities are processed with pfType = NO_USE_POST_FILTER (:474-476), whose
        # gain (stored in wp1, not applied) follows the |.| branch
        applied = (n - 1) >= min_frames and pf_type != NO_USE_POST_FILTER
        num = np.maximum(tot.real, 0.0) if (applied and (pf_type & TYPE_ZELINSKI1_REAL)) else np.abs(tot)
        with np.errstate(divide="ignore", invalid="ignore"):
            W = (num / Psi.sum(axis=1)) * (2.0 / (C - 1.0))
        W = np.where(W >= 1.0, 1.0, W)
        W = np.where(W < SPECTRAL_FLOOR, SPECTRAL_FLOOR, W)
        Wall[n] = W
        if applied:
            Ypf[n, :B] = W * Y[n, :B]
            Ypf[n, B:] = np.conj(Ypf[n, 1:M // 2][::-1])
    return Ypf, Wall


def chain_zelinski(pcm, h, g, geo: BankGeometry, W, ta, alpha=0.6, pf_type=TYPE_ZELINSKI1_ABS, min_frames=0, gain=1):
    """analysis -> beamform (weights W) -> Zelinski post-filter (manifold ta) -> synthesis.
    Returns (X, Y, Ypf, Wpf, out)."""
    pcm = np.asarray(pcm)
    C = pcm.shape[1]
    X = np.stack([analysis(pcm[:, c], h, geo) for c in range(C)], axis=1)
    Y = beamform(X, W)
    Ypf, Wpf = zelinski_postfilter(X, Y, ta, alpha, pf_type, min_frames)
    out = synthesis(Ypf, g, geo, gain)
    return X, Y, Ypf, Wpf, out.reshape(-1)


# --------------------------------------------------------------------------- compiled reference (oracle/_ref)
class _ChainCfg(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int), ("m", ctypes.c_int), ("r", ctypes.c_int), ("dct", ctypes.c_int), ("C", ctypes.c_int),
        ("fs", ctypes.c_double), ("mode", ctypes.c_int), ("inverse_kind", ctypes.c_int), ("noise_model", ctypes.c_int),
        ("dThreshold", ctypes.c_double), ("diag_load", ctypes.c_float), ("divide_mu", ctypes.c_float),
        ("sspeed", ctypes.c_double), ("gain", ctypes.c_int),
    ]


def _dp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class CompiledReference:
    """ctypes view of oracle/_ref/libbtk_ref.so (the reference's own code; see oracle/ref_driver.cc)."""

    def __init__(self, path: str | None = None):
        here = os.path.dirname(os.path.abspath(__file__))
        self.path = path or os.path.join(here, "_ref", "libbtk_ref.so")
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path)
        self.lib = ctypes.CDLL(self.path)
        L = self.lib
        vp, cl, ci, cd = ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_double
        L.btkref_analysis.restype = cl
        L.btkref_analysis.argtypes = [vp, cl, vp, ci, ci, ci, ci, vp, cl]
        L.btkref_synthesis.restype = cl
        L.btkref_synthesis.argtypes = [vp, cl, vp, ci, ci, ci, ci, ci, vp, cl]
        L.btkref_chain.restype = cl
        L.btkref_chain.argtypes = [ctypes.POINTER(_ChainCfg), vp, cl, vp, vp, vp, vp, vp, vp, vp, cl, vp, cl,
                                   ctypes.POINTER(cl), vp]
        if hasattr(L, "btkref_chain_zelinski"):
            L.btkref_chain_zelinski.restype = cl
            L.btkref_chain_zelinski.argtypes = [ctypes.POINTER(_ChainCfg), vp, cl, vp, vp, vp, cd, ci, ci, vp, vp, cl, vp,
                                                cl, ctypes.POINTER(cl)]
        if hasattr(L, "btkref_chain_gsc"):
            L.btkref_chain_gsc.restype = cl
            L.btkref_chain_gsc.argtypes = [ctypes.POINTER(_ChainCfg), vp, cl, vp, vp, vp, vp, ci, vp, cl, vp, cl,
                                           ctypes.POINTER(cl), vp, vp]
        if hasattr(L, "btkref_design_dehaan"):
            L.btkref_design_dehaan.restype = ci
            L.btkref_design_dehaan.argtypes = [ci, ci, ci, cd, cd, cd, vp, vp, vp, vp]
        L.btkref_spectral_matrix.restype = cl
        L.btkref_spectral_matrix.argtypes = [vp, cl, ci, vp, ci, ci, ci, ci, cd, vp]
        L.btkref_error_probe.restype = ci
        L.btkref_error_probe.argtypes = [ci]
        if hasattr(L, "btkref_iterative_sample"):
            L.btkref_iterative_sample.restype = cl
            L.btkref_iterative_sample.argtypes = [vp, cl, ci, ci, ci, ci, ci, vp, cl, ctypes.POINTER(cl)]
            L.btkref_conversion24.restype = cl
            L.btkref_conversion24.argtypes = [vp, cl, ci, vp]
            L.btkref_channel_extraction.restype = cl
            L.btkref_channel_extraction.argtypes = [vp, cl, ci, ci, ci, vp]
        if hasattr(L, "btkref_design_nyquist"):
            L.btkref_design_nyquist.restype = ci
            L.btkref_design_nyquist.argtypes = [ci, ci, ci, cd, cd, vp, vp, vp, vp]
        if hasattr(L, "btkref_null_weights"):
            L.btkref_null_weights.restype = ci
            L.btkref_null_weights.argtypes = [cd, vp, vp, ci, ci, ci, vp, vp]
            L.btkref_calc_all_delays.restype = ci
            L.btkref_calc_all_delays.argtypes = [cd, cd, cd, vp, ci, vp]
            L.btkref_calc_delays_polar2.restype = ci
            L.btkref_calc_delays_polar2.argtypes = [ctypes.c_float, ctypes.c_float, vp, ci, vp]

    @staticmethod
    def available(path: str | None = None) -> bool:
        here = os.path.dirname(os.path.abspath(__file__))
        return os.path.exists(path or os.path.join(here, "_ref", "libbtk_ref.so"))

    def analysis(self, x, h, geo: BankGeometry) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        h = np.ascontiguousarray(h, dtype=np.float64)
        cap = geo.analysis_frames(x.shape[0]) + 4
        X = np.zeros((cap, geo.M, 2), dtype=np.float64)
        n = self.lib.btkref_analysis(_dp(x), x.shape[0], _dp(h), geo.M, geo.m, geo.r, geo.dct, _dp(X), cap)
        if n < 0 or n > cap:
            raise RuntimeError(f"btkref_analysis returned {n}")
        return X[:n].view(np.complex128)[..., 0]

    def synthesis(self, Y, g, geo: BankGeometry, gain: int = 1) -> np.ndarray:
        Y = np.ascontiguousarray(Y, dtype=np.complex128)
        g = np.ascontiguousarray(g, dtype=np.float64)
        cap = Y.shape[0] + 4
        out = np.zeros((cap, geo.D), dtype=np.float32)
        n = self.lib.btkref_synthesis(_dp(Y), Y.shape[0], _dp(g), geo.M, geo.m, geo.r, geo.dct, gain, _dp(out), cap)
        if n < 0 or n > cap:
            raise RuntimeError(f"btkref_synthesis returned {n}")
        return out[:n]

    def chain(self, pcm, h, g, geo: BankGeometry, delays, fs=16000.0, mode="ds", Rn=None, micpos=None,
              diag_load=0.0, divide_mu=-1.0, dThreshold=1e-8, inverse="double", want_snap=True, want_Y=True,
              gain=1, sspeed=SSPEED):
        """Run the reference chain.  Returns dict(X=[F][C][M], Y=[F][M], out=[nblk*D], W=[B][C], frames, out_frames)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        T, C = pcm.shape
        h = np.ascontiguousarray(h, dtype=np.float64)
        g = None if g is None else np.ascontiguousarray(g, dtype=np.float64)
        delays = np.ascontiguousarray(delays, dtype=np.float64)
        cfg = _ChainCfg(geo.M, geo.m, geo.r, geo.dct, C, fs, 1 if mode == "mvdr" else 0,
                        1 if inverse == "double" else 0, 1 if (micpos is not None) else 0, dThreshold,
                        diag_load, divide_mu, sspeed, gain)
        cap = geo.analysis_frames(T) + 4
        snap = np.zeros((cap, geo.M, C, 2), dtype=np.float64) if want_snap else None
        Y = np.zeros((cap, geo.M, 2), dtype=np.float64) if want_Y else None
        cap_out = geo.nblk(T) + 4
        out = np.zeros((cap_out, geo.D), dtype=np.float32) if g is not None else None
        W = np.zeros((geo.B, C, 2), dtype=np.float64)
        Rn_c = None if Rn is None else np.ascontiguousarray(Rn, dtype=np.complex128)
        mp = None if micpos is None else np.ascontiguousarray(micpos, dtype=np.float64)
        nout = ctypes.c_long(0)
        n = self.lib.btkref_chain(ctypes.byref(cfg), _dp(pcm), T, _dp(h), _dp(g), _dp(delays), _dp(Rn_c), _dp(mp),
                                  _dp(snap), _dp(Y), cap, _dp(out), cap_out, ctypes.byref(nout), _dp(W))
        if n < 0 or n > cap:
            raise RuntimeError(f"btkref_chain returned {n}")
        res = {"frames": int(n), "out_frames": int(nout.value), "W": W.view(np.complex128)[..., 0]}
        if want_snap:
            res["X"] = np.ascontiguousarray(snap[:n].view(np.complex128)[..., 0].transpose(0, 2, 1))  # [F][C][M]
        if want_Y:
            res["Y"] = Y[:n].view(np.complex128)[..., 0]
        if g is not None:
            res["out"] = out[: nout.value].reshape(-1)
        return res

    def chain_zelinski(self, pcm, h, g, geo: BankGeometry, delays, alpha=0.6, pf_type=2, min_frames=0, fs=16000.0, gain=1):
        """The reference's SubbandDS -> ZelinskiPostFilter -> synthesis chain.  Returns dict(Ypf [F][M], Wpf [F][M], out)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        T, C = pcm.shape
        h = np.ascontiguousarray(h, dtype=np.float64)
        g = None if g is None else np.ascontiguousarray(g, dtype=np.float64)
        delays = np.ascontiguousarray(delays, dtype=np.float64)
        cfg = _ChainCfg(geo.M, geo.m, geo.r, geo.dct, C, fs, 0, 1, 0, 1e-8, 0.0, -1.0, SSPEED, gain)
        cap = geo.analysis_frames(T) + 4
        Ypf = np.zeros((cap, geo.M, 2), dtype=np.float64)
        Wpf = np.zeros((cap, geo.M), dtype=np.float64)
        cap_out = geo.nblk(T) + 4
        out = np.zeros((cap_out, geo.D), dtype=np.float32) if g is not None else None
        nout = ctypes.c_long(0)
        n = self.lib.btkref_chain_zelinski(ctypes.byref(cfg), _dp(pcm), T, _dp(h), _dp(g), _dp(delays), alpha, pf_type,
                                           min_frames, _dp(Ypf), _dp(Wpf), cap, _dp(out), cap_out, ctypes.byref(nout))
        if n < 0 or n > cap:
            raise RuntimeError(f"btkref_chain_zelinski returned {n}")
        res = {"frames": int(n), "Ypf": Ypf[:n].view(np.complex128)[..., 0], "Wpf": Wpf[:n]}
        if g is not None:
            res["out"] = out[: nout.value].reshape(-1)
        return res

    def chain_gsc(self, pcm, h, g, geo: BankGeometry, delays, wa, normalize=False, fs=16000.0, gain=1):
        """The reference's SubbandGSC (calcGSCWeights + setActiveWeights_f per bin) -> synthesis.
        wa [B][C-1] complex.  Returns dict(Y [F][M], out, Bm [B][C][C-1], wq [B][C])."""
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        T, C = pcm.shape
        h = np.ascontiguousarray(h, dtype=np.float64)
        g = None if g is None else np.ascontiguousarray(g, dtype=np.float64)
        delays = np.ascontiguousarray(delays, dtype=np.float64)
        wa = np.ascontiguousarray(wa, dtype=np.complex128)
        cfg = _ChainCfg(geo.M, geo.m, geo.r, geo.dct, C, fs, 2, 1, 0, 1e-8, 0.0, -1.0, SSPEED, gain)
        cap = geo.analysis_frames(T) + 4
        Y = np.zeros((cap, geo.M, 2), dtype=np.float64)
        cap_out = geo.nblk(T) + 4
        out = np.zeros((cap_out, geo.D), dtype=np.float32) if g is not None else None
        Bm = np.zeros((geo.B, C, C - 1, 2), dtype=np.float64)
        wq = np.zeros((geo.B, C, 2), dtype=np.float64)
        nout = ctypes.c_long(0)
        n = self.lib.btkref_chain_gsc(ctypes.byref(cfg), _dp(pcm), T, _dp(h), _dp(g), _dp(delays), _dp(wa),
                                      1 if normalize else 0, _dp(Y), cap, _dp(out), cap_out, ctypes.byref(nout), _dp(Bm),
                                      _dp(wq))
        if n < 0 or n > cap:
            raise RuntimeError(f"btkref_chain_gsc returned {n}")
        res = {"frames": int(n), "Y": Y[:n].view(np.complex128)[..., 0], "Bm": Bm.view(np.complex128)[..., 0],
               "wq": wq.view(np.complex128)[..., 0]}
        if g is not None:
            res["out"] = out[: nout.value].reshape(-1)
        return res

    def design_dehaan(self, M, m, r, wp_factor=1.0, v=1.0, tol=1e-7):
        """The reference's AnalysisOversampledDFTDesign + SynthesisOversampledDFTDesign.  Returns (h, g, err_h, err_g)."""
        L = M * m
        h, g, eh, eg = np.zeros(L), np.zeros(L), np.zeros(3), np.zeros(3)
        rc = self.lib.btkref_design_dehaan(M, m, r, wp_factor, v, tol, _dp(h), _dp(g), _dp(eh), _dp(eg))
        if rc != 0:
            raise RuntimeError(f"btkref_design_dehaan returned {rc}")
        return h, g, eh, eg

    def iterative_sample(self, pcm, samplerate, blockLen, cfrom=0, cto=-1):
        """IterativeSampleFeature::next of the reference (extracted from feature/feature.cc at build time), one node per
        channel pulled in lock step: (blocks [n][C][blockLen], samples read)."""
        x = np.ascontiguousarray(pcm, dtype=np.float32)
        T, C = x.shape
        blockN = 30 * samplerate // blockLen + 1
        cap = (T // (blockN * blockLen) + 2) * blockN
        out = np.zeros((cap, C, blockLen), np.float32)
        ttl = ctypes.c_long(0)
        n = self.lib.btkref_iterative_sample(_dp(x), T, samplerate, C, blockLen, cfrom, cto, _dp(out), cap, ctypes.byref(ttl))
        if n < 0 or n > cap:
            raise RuntimeError(f"btkref_iterative_sample returned {n}")
        return out[:n], int(ttl.value)

    def conversion24(self, raw, block):
        """Conversion24bit2Float::next of the reference over packed big-endian 24-bit bytes (uint8 [n][3])."""
        b = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1)
        out = np.zeros(b.size // 3, np.float32)
        n = self.lib.btkref_conversion24(_dp(b), b.size, block, _dp(out))
        return out[:n]

    def channel_extraction(self, data, chX, chN, block):
        x = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
        out = np.zeros(x.size // chN, np.float32)
        n = self.lib.btkref_channel_extraction(_dp(x), x.size, chX, chN, block, _dp(out))
        return out[:n]

    def design_nyquist(self, M, m, r, wp_factor=1.0, tol=1e-7):
        """AnalysisNyquistMDesign then SynthesisNyquistMDesign of the compiled reference: (h, g)."""
        L = M * m
        h, g = np.zeros(L), np.zeros(L)
        eh, eg = np.zeros(3), np.zeros(3)
        if self.lib.btkref_design_nyquist(M, m, r, wp_factor, tol, _dp(h), _dp(g), _dp(eh), _dp(eg)) != 0:
            raise RuntimeError("btkref_design_nyquist failed")
        return h, g

    def spectral_matrix(self, pcm, h, geo: BankGeometry, mu=0.95) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        T, C = pcm.shape
        h = np.ascontiguousarray(h, dtype=np.float64)
        R = np.zeros((geo.M, C, C, 2), dtype=np.float64)
        n = self.lib.btkref_spectral_matrix(_dp(pcm), T, C, _dp(h), geo.M, geo.m, geo.r, geo.dct, mu, _dp(R))
        if n < 0:
            raise RuntimeError(f"btkref_spectral_matrix returned {n}")
        return R.view(np.complex128)[..., 0]

    def null_weights(self, delaysT, delaysJ, fs: float, M: int):
        """SubbandDS::calcArrayManifoldVectors2 / N of the compiled reference: (wq [B][C], array manifold [B][C])."""
        dT = np.ascontiguousarray(delaysT, dtype=np.float64)
        dJ = np.ascontiguousarray(np.atleast_2d(delaysJ), dtype=np.float64)
        C, NC, B = dT.shape[0], dJ.shape[0] + 1, M // 2 + 1
        w = np.zeros((B, C), dtype=np.complex128)
        ta = np.zeros((B, C), dtype=np.complex128)
        rc = self.lib.btkref_null_weights(fs, _dp(dT), _dp(dJ), M, C, NC, _dp(w), _dp(ta))
        if rc != 0:
            raise RuntimeError("btkref_null_weights failed")
        return w, ta

    def all_delays(self, micpos, x=0.0, y=0.0, z=0.0) -> np.ndarray:
        mp = np.ascontiguousarray(micpos, dtype=np.float64)
        d = np.zeros(mp.shape[0], dtype=np.float64)
        self.lib.btkref_calc_all_delays(x, y, z, _dp(mp), mp.shape[0], _dp(d))
        return d

    def delays_polar2(self, azimuth, elevation, micpos) -> np.ndarray:
        mp = np.ascontiguousarray(micpos, dtype=np.float64)
        d = np.zeros(mp.shape[0], dtype=np.float64)
        self.lib.btkref_calc_delays_polar2(azimuth, elevation, _dp(mp), mp.shape[0], _dp(d))
        return d

    def error_probe(self, which: int) -> int:
        return int(self.lib.btkref_error_probe(which))


# --------------------------------------------------------------------------- metrics
def ingest_s16(raw: np.ndarray) -> np.ndarray:
    """16-bit PCM -> float32 without normalisation: what sf_readf_float hands IterativeSampleFeature when
    SFC_SET_NORM_FLOAT is off (feature/feature.cc:273, 849, 868-896).  Same element order."""
    return np.asarray(raw, dtype=np.int16).astype(np.float32)


def ingest_s24be(raw: np.ndarray) -> np.ndarray:
    """Packed big-endian 24-bit -> float32, a loop restatement of Conversion24bit2Float::next
    (feature/feature.cc:190-217): byte 0 is the most significant one and carries the sign; Mark-III/IV data
    frames are 64 channels x 3 bytes (driver/mk4_common.h:50-54).  raw: uint8 [..., 3]."""
    b = np.asarray(raw, dtype=np.uint8).astype(np.int64)
    v = (b[..., 0] << 16) | (b[..., 1] << 8) | b[..., 2]
    v = np.where(b[..., 0] & 128, v - (1 << 24), v)
    return v.astype(np.float32)


def iterative_sample_blocks(pcm: np.ndarray, samplerate: int, blockLen: int, cfrom: int = 0, cto: int = -1, interval: int = 30):
    """IterativeSampleFeature (feature/feature.cc:803-896) for every channel of an interleaved recording pulled in lock step:
    the node of channel `firstChanX` refills a shared buffer of _blockN = interval * samplerate / blockLen + 1 blocks
    (zeroed first, :886-888) whenever its block counter wraps; a short read sets _last and the stream ends at the NEXT wrap
    (:880-884) -- so the stream is a whole number of 30-s buffers long, zero padded.  With cto > 0 the end test is
    cur * blockLen > cto - cfrom, evaluated only at a wrap.  pcm: [T][C].  Returns (blocks [n][C][blockLen], samples read)."""
    pcm = np.asarray(pcm, dtype=np.float32)
    T, C = pcm.shape
    blockN = interval * samplerate // blockLen + 1
    sampleN = blockN * blockLen
    pos, ctor, cur, last, ttl = cfrom, cto - cfrom, 0, False, 0
    buf = np.zeros((sampleN, C), np.float32)
    out = []
    while True:
        cf = cur % blockN
        if cf == 0:
            if last or (ctor > 0 and cur * blockLen > ctor):
                break
            buf[:] = 0
            n = max(0, min(sampleN, T - pos))
            buf[:n] = pcm[pos:pos + n]
            pos += n
            ttl += n
            if n < sampleN:
                last = True
        out.append(buf[cf * blockLen:(cf + 1) * blockLen].T.copy())
        cur += 1
    return (np.stack(out) if out else np.zeros((0, C, blockLen), np.float32)), ttl


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    """||a - b||_2 / ||b||_2 (b is the reference)."""
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))


def snr_db(test: np.ndarray, ref: np.ndarray) -> float:
    """10 log10(||ref||^2 / ||test - ref||^2)."""
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(test, dtype=np.float64) - ref
    pe = float((err**2).sum())
    ps = float((ref**2).sum())
    if pe == 0:
        return float("inf")
    return 10.0 * np.log10(ps / pe)
