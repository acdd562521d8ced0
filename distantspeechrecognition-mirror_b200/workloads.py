"""Synthetic array recordings and array geometries for the BASELINE.json configs.

Host-side numpy only (input generation is not part of the hot path).  All signals are float32 in the
int16 range, like SampleFeature's un-normalised samples (reference btk/feature/feature.cc:273), laid out
interleaved [T][C] like IterativeSampleFeature's source buffer (btk/feature/feature.cc:868-896).
"""
from __future__ import annotations

import numpy as np

SSPEED_MM_S = 343740.0  # reference btk/beamformer/beamformer.h:47
FS = 16000.0


def circular_array(C: int = 8, radius_mm: float = 100.0) -> np.ndarray:
    a = 2.0 * np.pi * np.arange(C) / C
    return np.stack([radius_mm * np.cos(a), radius_mm * np.sin(a), np.zeros(C)], axis=1)


def linear_array(C: int, pitch_mm: float) -> np.ndarray:
    x = pitch_mm * (np.arange(C) - (C - 1) / 2.0)
    return np.stack([x, np.zeros(C), np.zeros(C)], axis=1)


def farfield_delays(micpos_mm: np.ndarray, azimuth: float, elevation: float) -> np.ndarray:
    """tau_c = (c . p_c)/343740 with c = -(sin(el)cos(az), sin(el)sin(az), cos(el))
    (reference btk/src/superdirectiveBeamformer.cc:118-137)."""
    c = -np.array([np.sin(elevation) * np.cos(azimuth), np.sin(elevation) * np.sin(azimuth), np.cos(elevation)])
    return (np.asarray(micpos_mm, dtype=np.float64)[:, :3] @ c) / SSPEED_MM_S


def chirp(T: int, f0: float = 100.0, f1: float = 7000.0, amp: float = 8000.0, fs: float = FS) -> np.ndarray:
    t = np.arange(T) / fs
    dur = max(T / fs, 1e-9)
    return amp * np.sin(2.0 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / dur))


def array_recording(T: int, delays_s: np.ndarray, seed: int, noise_sigma: float = 100.0,
                    source_amp: float = 8000.0, fs: float = FS) -> np.ndarray:
    """Chirp source seen by every microphone with its own (fractional) delay, applied exactly in the
    frequency domain, plus i.i.d. Gaussian sensor noise.  Returns float32 [T][C]."""
    rng = np.random.default_rng(seed)
    C = len(delays_s)
    s = chirp(T, amp=source_amp, fs=fs)
    nfft = 1 << int(np.ceil(np.log2(T + 64)))
    S = np.fft.rfft(s, nfft)
    f = np.fft.rfftfreq(nfft, 1.0 / fs)
    out = np.empty((T, C), dtype=np.float32)
    for c in range(C):
        # channel c hears s(t - tau_c) (tau_c = -(u . p_c)/v is negative for microphones nearer the
        # source), so Y = sum_c conj(wq_c) X_c with wq_c = exp(-j w tau_c)/C (beamformer.cc:566-574) re-aligns it
        sc = np.fft.irfft(S * np.exp(-2j * np.pi * f * delays_s[c]), nfft)[:T]
        out[:, c] = (sc + noise_sigma * rng.standard_normal(T)).astype(np.float32)
    return out


def noise_recording(T: int, C: int, seed: int, sigma: float = 1000.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (sigma * rng.standard_normal((T, C))).astype(np.float32)


_DESIGNED = {}


def designed_prototype(M: int, m: int, r: int) -> tuple[np.ndarray, np.ndarray]:
    """Analysis/synthesis prototypes for an (M, m, r) without a reference fixture, DESIGNED on the device with the
    reference's de Haan procedure (btkb200_design_analysis_prototype / _synthesis_prototype: h = pinv(A + C) b,
    g = pinv(E + v P) f, modulated/prototypeDesign.cc:611-951, v = 1, wpFactor = 1, tolerance 1e-7).  Cached per process;
    needs a CUDA device (the stand-in below is what the CPU-only tests use)."""
    key = (int(M), int(m), int(r))
    if key not in _DESIGNED:
        from . import _capi
        h, _ = _capi.design_analysis_prototype(M, m, r)
        g, _ = _capi.design_synthesis_prototype(h, M, m, r, 1.0)
        _DESIGNED[key] = (h, g)
    return _DESIGNED[key]


def kaiser_prototype(M: int, m: int, r: int, beta: float = 8.0) -> tuple[np.ndarray, np.ndarray]:
    """Stand-in analysis/synthesis prototypes for (M, m, r) without a reference fixture: Kaiser-windowed
    sinc low-pass with cut-off pi/M, unit DC gain for h and D-scaled for g.  Throughput does not depend on
    tap values and parity is always judged against the oracle run with the SAME taps."""
    N = M * m
    n = np.arange(N) - (N - 1) / 2.0
    h = np.sinc(n / M) * np.kaiser(N, beta)
    h = h / h.sum()
    D = M >> r
    return h.astype(np.float64), (h * D).astype(np.float64)
