"""ctypes binding of include/btkb200.h (libbtkb200.so, built in-tree by csrc/Makefile).

This is plumbing only: every number is produced by the sm_100a kernels behind the C ABI.  There is no
CPU fallback -- if the library is missing or no GPU is visible the calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_long, c_longlong, c_size_t, c_uint, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BTKB200_LIB") or os.path.join(_HERE, "libbtkb200.so")   # override: A/B builds while tuning

OK, EINVAL, ESTATE, ECUDA, ENOMEM, EUNSUPPORTED = 0, 1, 2, 3, 4, 5
PCM_F32, PCM_S16, PCM_S24BE = 0, 1, 2      # raw PCM formats of chain_batch_pcm / convert_pcm (include/btkb200.h)
TUNE_CHAIN_WS, TUNE_CLUSTER = 1, 2         # btkb200_plan_tune knobs (include/btkb200.h)

# every symbol include/btkb200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "btkb200_device_count", "btkb200_version", "btkb200_last_error", "btkb200_plan_create", "btkb200_plan_destroy",
    "btkb200_plan_info", "btkb200_nblk", "btkb200_analysis_frames", "btkb200_synthesis_frames", "btkb200_chain_frames",
    "btkb200_set_ds_weights", "btkb200_set_weights", "btkb200_get_weights", "btkb200_get_manifold",
    "btkb200_set_covariance", "btkb200_get_covariance", "btkb200_set_diffuse_noise_model", "btkb200_diag_load",
    "btkb200_diag_load_bin", "btkb200_divide_nondiagonal", "btkb200_solve_mvdr", "btkb200_analysis",
    "btkb200_beamform", "btkb200_synthesis", "btkb200_covariance", "btkb200_estimate_covariance", "btkb200_chain",
    "btkb200_chain_batch", "btkb200_mvdr_chain_batch", "btkb200_chain_batch_pcm", "btkb200_convert_pcm", "btkb200_beamform_zelinski",
    "btkb200_chain_zelinski", "btkb200_chain_zelinski_batch", "btkb200_beamform_zelinski_dev", "btkb200_gsc_calc_weights",
    "btkb200_design_analysis_prototype", "btkb200_design_synthesis_prototype", "btkb200_gsc_set_active_weights", "btkb200_gsc_zero_active_weights", "btkb200_gsc_get_blocking_matrix", "btkb200_gsc_apply",
    "btkb200_chain_batch_multi", "btkb200_chain_batch_dev", "btkb200_analysis_dev", "btkb200_beamform_dev",
    "btkb200_synthesis_dev", "btkb200_launch_count", "btkb200_sync", "btkb200_host_alloc", "btkb200_host_free",
    "btkb200_plan_tune", "btkb200_plan_tuning", "btkb200_set_null_weights", "btkb200_calc_delays_polar",
    "btkb200_calc_all_delays", "btkb200_get_array_manifold", "btkb200_design_analysis_nyquist",
    "btkb200_design_synthesis_nyquist",
]


class MvdrAdapt(ctypes.Structure):
    _fields_ = [("forget", c_double), ("last_frame", c_long), ("conjugate", c_int), ("load_abs", c_double),
                ("load_rel", c_double), ("dThreshold", c_double)]


class Info(ctypes.Structure):
    _fields_ = [(n, c_uint) for n in ("M", "m", "r", "R", "D", "N", "B", "C", "dct", "pd_analysis", "pd_synthesis",
                                      "laN")] + [("device", c_int), ("has_weights", c_int)]


class BtkError(RuntimeError):
    """A non-zero status from the C ABI.  ``code`` is one of EINVAL/ESTATE/ECUDA/ENOMEM/EUNSUPPORTED."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"btkb200 status {code}: {msg}")
        self.code = code
        self.msg = msg


_lib = None


def lib() -> ctypes.CDLL:
    """Load libbtkb200.so (once).  Fails loudly when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                          "(or __graft_entry__.build()).  There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp = c_void_p
    L.btkb200_device_count.restype = c_int
    L.btkb200_version.restype = c_char_p
    L.btkb200_last_error.restype = c_char_p
    L.btkb200_last_error.argtypes = [vp]
    L.btkb200_plan_create.restype = c_int
    L.btkb200_plan_create.argtypes = [POINTER(vp), c_uint, c_uint, c_uint, c_uint, c_uint, vp, vp, c_int, c_int]
    L.btkb200_plan_destroy.restype = None
    L.btkb200_plan_destroy.argtypes = [vp]
    L.btkb200_plan_info.argtypes = [vp, POINTER(Info)]
    for f in ("btkb200_nblk", "btkb200_analysis_frames", "btkb200_synthesis_frames", "btkb200_chain_frames"):
        getattr(L, f).restype = c_long
        getattr(L, f).argtypes = [vp, c_long]
    L.btkb200_set_ds_weights.argtypes = [vp, c_double, vp, c_uint]
    L.btkb200_set_null_weights.argtypes = [vp, c_double, vp, c_uint, vp, c_uint]
    L.btkb200_calc_delays_polar.argtypes = [c_float, c_float, vp, c_uint, vp]
    L.btkb200_calc_all_delays.argtypes = [c_double, c_double, c_double, vp, c_uint, vp]
    L.btkb200_get_array_manifold.argtypes = [vp, vp]
    L.btkb200_set_weights.argtypes = [vp, vp]
    L.btkb200_get_weights.argtypes = [vp, vp]
    L.btkb200_get_manifold.argtypes = [vp, vp]
    L.btkb200_set_covariance.argtypes = [vp, c_uint, vp, c_uint, c_uint]
    L.btkb200_get_covariance.argtypes = [vp, c_uint, vp]
    L.btkb200_set_diffuse_noise_model.argtypes = [vp, vp, c_uint, c_double, c_double]
    L.btkb200_diag_load.argtypes = [vp, c_float]
    L.btkb200_diag_load_bin.argtypes = [vp, c_uint, c_float]
    L.btkb200_divide_nondiagonal.argtypes = [vp, c_float]
    L.btkb200_solve_mvdr.argtypes = [vp, c_double, c_double, POINTER(c_int)]
    L.btkb200_analysis.argtypes = [vp, vp, c_long, vp, POINTER(c_long)]
    L.btkb200_beamform.argtypes = [vp, vp, c_long, vp]
    L.btkb200_synthesis.argtypes = [vp, vp, c_long, vp, POINTER(c_long)]
    L.btkb200_covariance.argtypes = [vp, vp, c_long, vp, c_int, vp]
    L.btkb200_estimate_covariance.argtypes = [vp, vp, c_long, c_double, c_long, c_int]
    L.btkb200_chain.argtypes = [vp, vp, c_long, vp]
    L.btkb200_chain_batch.argtypes = [vp, POINTER(vp), POINTER(c_long), c_int, POINTER(vp)]
    L.btkb200_mvdr_chain_batch.argtypes = [vp, POINTER(vp), POINTER(c_long), c_int, POINTER(MvdrAdapt), POINTER(vp), vp]
    L.btkb200_chain_batch_pcm.argtypes = [vp, POINTER(vp), c_int, POINTER(c_long), c_int, POINTER(vp)]
    L.btkb200_convert_pcm.argtypes = [vp, vp, c_int, c_long, vp]
    L.btkb200_beamform_zelinski.argtypes = [vp, vp, c_long, c_double, c_int, c_int, vp, vp]
    L.btkb200_chain_zelinski.argtypes = [vp, vp, c_long, c_double, c_int, c_int, vp]
    L.btkb200_chain_zelinski_batch.argtypes = [vp, POINTER(vp), POINTER(c_long), c_int, c_double, c_int, c_int, POINTER(vp)]
    L.btkb200_beamform_zelinski_dev.argtypes = [vp, vp, c_long, c_double, c_int, c_int, vp, vp, vp]
    L.btkb200_design_analysis_prototype.argtypes = [c_uint, c_uint, c_uint, c_double, c_int, c_double, c_int, vp, vp]
    L.btkb200_design_synthesis_prototype.argtypes = [vp, c_uint, c_uint, c_uint, c_double, c_double, c_int, c_double, c_int, vp, vp]
    L.btkb200_gsc_calc_weights.argtypes = [vp, c_double, vp, c_uint]
    L.btkb200_gsc_set_active_weights.argtypes = [vp, c_uint, vp, c_uint]
    L.btkb200_gsc_zero_active_weights.argtypes = [vp]
    L.btkb200_gsc_get_blocking_matrix.argtypes = [vp, c_uint, vp]
    L.btkb200_gsc_apply.argtypes = [vp, c_int]
    L.btkb200_chain_batch_multi.argtypes = [POINTER(vp), c_int, POINTER(vp), POINTER(c_long), c_int, POINTER(vp)]
    L.btkb200_chain_batch_dev.argtypes = [vp, vp, POINTER(c_longlong), POINTER(c_longlong), POINTER(c_longlong), c_int,
                                          vp, vp]
    L.btkb200_analysis_dev.argtypes = [vp, vp, c_long, vp, vp]
    L.btkb200_beamform_dev.argtypes = [vp, vp, c_long, vp, vp]
    L.btkb200_synthesis_dev.argtypes = [vp, vp, c_long, vp, vp]
    L.btkb200_plan_tune.argtypes = [vp, c_int, c_int]
    L.btkb200_plan_tuning.argtypes = [vp, c_int]
    L.btkb200_launch_count.restype = c_long
    L.btkb200_launch_count.argtypes = [vp]
    L.btkb200_sync.argtypes = [vp]
    L.btkb200_host_alloc.restype = vp
    L.btkb200_host_alloc.argtypes = [c_size_t]
    L.btkb200_host_free.restype = None
    L.btkb200_host_free.argtypes = [vp]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


class Plan:
    """Owner of one ``btkb200_plan`` (geometry + prototypes + beamformer weights on one device)."""

    def __init__(self, M: int, m: int, r: int, C: int, h=None, g=None, dct: int = 0, gain: int = 1, device: int = 0):
        self._L = lib()
        self._h = c_void_p()
        hh = None if h is None else np.ascontiguousarray(h, dtype=np.float64)
        gg = None if g is None else np.ascontiguousarray(g, dtype=np.float64)
        for name, a in (("analysis", hh), ("synthesis", gg)):
            if a is not None and a.size != M * m:
                # OverSampledDFTFilterBank ctor: jconsistency_error (reference modulated/modulated.cc:269-271)
                raise BtkError(EINVAL, f"Prototype sizes do not match ({a.size} vs. {M * m}).")
        rc = self._L.btkb200_plan_create(ctypes.byref(self._h), M, m, r, dct, C, _p(hh), _p(gg), gain, device)
        if rc != OK:
            raise BtkError(rc, (self._L.btkb200_last_error(None) or b"").decode())
        inf = Info()
        self._L.btkb200_plan_info(self._h, ctypes.byref(inf))
        self.info = inf
        self.M, self.m, self.r, self.R, self.D, self.N, self.B, self.C = (inf.M, inf.m, inf.r, inf.R, inf.D, inf.N,
                                                                          inf.B, inf.C)
        self.pd_analysis, self.pd_synthesis, self.laN, self.device = inf.pd_analysis, inf.pd_synthesis, inf.laN, device

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.btkb200_plan_destroy(self._h)
            self._h = c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != OK:
            raise BtkError(rc, (self._L.btkb200_last_error(self._h) or b"").decode())

    # -- geometry
    def nblk(self, T: int) -> int:
        return int(self._L.btkb200_nblk(self._h, T))

    def chain_frames(self, T: int) -> int:
        return int(self._L.btkb200_chain_frames(self._h, T))

    def analysis_frames(self, T: int) -> int:
        return int(self._L.btkb200_analysis_frames(self._h, T))

    def synthesis_frames(self, F: int) -> int:
        return int(self._L.btkb200_synthesis_frames(self._h, F))

    # -- weights
    def has_weights(self) -> int:
        """0 none, 1 delay-and-sum, 2 user / MVDR (re-queried: the flag changes after construction)."""
        inf = Info()
        self._L.btkb200_plan_info(self._h, ctypes.byref(inf))
        return int(inf.has_weights)

    def set_ds_weights(self, fs: float, delays):
        d = np.ascontiguousarray(delays, dtype=np.float64)
        self._ck(self._L.btkb200_set_ds_weights(self._h, fs, _p(d), d.size))

    def set_null_weights(self, fs: float, delaysT, delaysJ):
        """SubbandDS::calcArrayManifoldVectors2 / N (beamformer.cc:1100-1121): delaysJ is [C] or [NC-1][C]."""
        dT = np.ascontiguousarray(delaysT, dtype=np.float64)
        dJ = np.ascontiguousarray(np.atleast_2d(np.asarray(delaysJ, dtype=np.float64)))
        if dJ.ndim != 2 or dJ.shape[1] != self.C:
            raise BtkError(EINVAL, f"interferer delays must be [NC-1][{self.C}], got {dJ.shape}")
        self._ck(self._L.btkb200_set_null_weights(self._h, fs, _p(dT), dT.size, _p(dJ), dJ.shape[0] + 1))

    def get_array_manifold(self) -> np.ndarray:
        w = np.zeros((self.B, self.C), dtype=np.complex128)
        self._ck(self._L.btkb200_get_array_manifold(self._h, _p(w)))
        return w

    def set_weights(self, W):
        w = np.ascontiguousarray(W, dtype=np.complex128)
        if w.shape != (self.B, self.C):
            raise BtkError(EINVAL, f"weights must be [{self.B}][{self.C}], got {w.shape}")
        self._ck(self._L.btkb200_set_weights(self._h, _p(w)))

    def get_weights(self) -> np.ndarray:
        w = np.zeros((self.B, self.C), dtype=np.complex128)
        self._ck(self._L.btkb200_get_weights(self._h, _p(w)))
        return w

    def get_manifold(self) -> np.ndarray:
        w = np.zeros((self.B, self.C), dtype=np.complex128)
        self._ck(self._L.btkb200_get_manifold(self._h, _p(w)))
        return w

    # -- MVDR
    def set_covariance(self, bin_: int, R):
        Rm = np.ascontiguousarray(R, dtype=np.complex128)
        if Rm.ndim != 2:
            raise BtkError(EINVAL, "covariance must be a matrix")
        self._ck(self._L.btkb200_set_covariance(self._h, bin_, _p(Rm), Rm.shape[0], Rm.shape[1]))

    def get_covariance(self, bin_: int) -> np.ndarray:
        R = np.zeros((self.C, self.C), dtype=np.complex128)
        self._ck(self._L.btkb200_get_covariance(self._h, bin_, _p(R)))
        return R

    def set_diffuse_noise_model(self, micpos_mm, fs: float, sspeed: float = 343740.0):
        mp = np.ascontiguousarray(micpos_mm, dtype=np.float64)
        if mp.ndim != 2 or mp.shape[1] < 3:
            raise BtkError(EINVAL, "The microphone positions should be described in the three dimensions")
        mp = np.ascontiguousarray(mp[:, :3])
        self._ck(self._L.btkb200_set_diffuse_noise_model(self._h, _p(mp), mp.shape[0], fs, sspeed))

    def diag_load(self, w: float, bin_: int | None = None):
        if bin_ is None:
            self._ck(self._L.btkb200_diag_load(self._h, w))
        else:
            self._ck(self._L.btkb200_diag_load_bin(self._h, bin_, w))

    def divide_nondiagonal(self, mu: float):
        self._ck(self._L.btkb200_divide_nondiagonal(self._h, mu))

    def solve_mvdr(self, fs: float = 16000.0, dThreshold: float = 1e-8) -> int:
        nfb = c_int(0)
        self._ck(self._L.btkb200_solve_mvdr(self._h, fs, dThreshold, ctypes.byref(nfb)))
        return int(nfb.value)

    # -- staged path (host numpy buffers)
    def analysis(self, pcm) -> np.ndarray:
        """pcm float32 [T][C] -> snapshots complex64 [F][B][C]."""
        x = np.ascontiguousarray(pcm, dtype=np.float32)
        if x.ndim == 1:
            x = x[:, None]
        if x.shape[1] != self.C:
            raise BtkError(EINVAL, f"pcm has {x.shape[1]} channels, plan has {self.C}")
        T = x.shape[0]
        F = self.analysis_frames(T)
        snap = np.empty((F, self.B, self.C), dtype=np.complex64)
        n = c_long(0)
        self._ck(self._L.btkb200_analysis(self._h, _p(x), T, _p(snap), ctypes.byref(n)))
        return snap

    def beamform(self, snap) -> np.ndarray:
        s = np.ascontiguousarray(snap, dtype=np.complex64)
        self._check_snap(s)
        F = s.shape[0]
        Y = np.empty((F, self.B), dtype=np.complex64)
        self._ck(self._L.btkb200_beamform(self._h, _p(s), F, _p(Y)))
        return Y

    def synthesis(self, Y) -> np.ndarray:
        y = np.ascontiguousarray(Y, dtype=np.complex64)
        if y.ndim != 2 or y.shape[1] != self.B:
            raise BtkError(EINVAL, f"beamformer output must be [F][{self.B}], got {y.shape}")
        F = y.shape[0]
        nout = self.synthesis_frames(F)
        out = np.empty(nout * self.D, dtype=np.float32)
        n = c_long(0)
        self._ck(self._L.btkb200_synthesis(self._h, _p(y), F, _p(out), ctypes.byref(n)))
        return out

    def covariance(self, snap, frame_weights, conjugate: bool = True) -> np.ndarray:
        s = np.ascontiguousarray(snap, dtype=np.complex64)
        self._check_snap(s)
        F = s.shape[0]
        w = np.ascontiguousarray(frame_weights, dtype=np.float64)
        if w.shape != (F,):
            raise BtkError(EINVAL, "one weight per frame expected")
        R = np.zeros((self.B, self.C, self.C), dtype=np.complex128)
        self._ck(self._L.btkb200_covariance(self._h, _p(s), F, _p(w), 1 if conjugate else 0, _p(R)))
        return R

    def estimate_covariance(self, pcm, forget: float = 0.99, last_frame: int = -1, conjugate: bool = True):
        """pcm [T][C] -> the plan's per-bin noise matrices (analysis + covariance on the device)."""
        x = np.ascontiguousarray(pcm, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.C:
            raise BtkError(EINVAL, f"pcm must be [T][{self.C}]")
        self._ck(self._L.btkb200_estimate_covariance(self._h, _p(x), x.shape[0], forget, last_frame, 1 if conjugate else 0))

    # -- SubbandGSC with fixed active weights (beamformer.cc:1296-1447)
    def gsc_calc_weights(self, fs: float, delays):
        d = np.ascontiguousarray(delays, dtype=np.float64)
        self._ck(self._L.btkb200_gsc_calc_weights(self._h, fs, _p(d), d.size))

    def gsc_set_active_weights(self, fbin: int, packed):
        w = np.ascontiguousarray(packed, dtype=np.float64).ravel()
        self._ck(self._L.btkb200_gsc_set_active_weights(self._h, fbin, _p(w), w.size))

    def gsc_zero_active_weights(self):
        self._ck(self._L.btkb200_gsc_zero_active_weights(self._h))

    def gsc_blocking_matrix(self, fbin: int) -> np.ndarray:
        Bm = np.zeros((self.C, self.C - 1), dtype=np.complex128)
        self._ck(self._L.btkb200_gsc_get_blocking_matrix(self._h, fbin, _p(Bm)))
        return Bm

    def gsc_apply(self, normalize: bool = False):
        self._ck(self._L.btkb200_gsc_apply(self._h, 1 if normalize else 0))

    # -- Zelinski post-filter (postfilter/postfilter.cc:30-222, 428-500)
    def beamform_zelinski(self, snap, alpha: float = 0.6, pf_type: int = 2, min_frames: int = 0):
        """snapshots [F][B][C] -> (post-filtered beamformer output [F][B] complex64, gains [F][B] float32)."""
        s = np.ascontiguousarray(snap, dtype=np.complex64)
        F = s.shape[0]
        if s.shape[1:] != (self.B, self.C):
            raise BtkError(EINVAL, f"snapshots must be [F][{self.B}][{self.C}]")
        Y = np.empty((F, self.B), dtype=np.complex64)
        W = np.empty((F, self.B), dtype=np.float32)
        self._ck(self._L.btkb200_beamform_zelinski(self._h, _p(s), F, alpha, pf_type, min_frames, _p(Y), _p(W)))
        return Y, W

    def chain_zelinski(self, pcm, alpha: float = 0.6, pf_type: int = 2, min_frames: int = 0) -> np.ndarray:
        """pcm [T][C] -> analysis -> beamformer -> Zelinski post-filter -> synthesis, all on the device."""
        x = np.ascontiguousarray(pcm, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.C:
            raise BtkError(EINVAL, f"pcm must be [T][{self.C}]")
        out = np.empty(self.chain_frames(x.shape[0]) * self.D, dtype=np.float32)
        self._ck(self._L.btkb200_chain_zelinski(self._h, _p(x), x.shape[0], alpha, pf_type, min_frames, _p(out)))
        return out

    def chain_zelinski_batch(self, pcms, alpha: float = 0.6, pf_type: int = 2, min_frames: int = 0) -> list:
        xs = [np.ascontiguousarray(x, dtype=np.float32) for x in pcms]
        outs = [np.empty(self.chain_frames(x.shape[0]) * self.D, dtype=np.float32) for x in xs]
        self.chain_zelinski_batch_into(xs, outs, alpha, pf_type, min_frames)
        return outs

    def chain_zelinski_batch_into(self, xs, outs, alpha: float = 0.6, pf_type: int = 2, min_frames: int = 0):
        """xs / outs: lists of C-contiguous float32 arrays (outs preallocated, ideally pinned); no allocation here."""
        n = len(xs)
        self._check_pcm_list(xs)
        self._check_out_list([x.shape[0] for x in xs], outs)
        pp = (c_void_p * n)(*[x.ctypes.data for x in xs])
        oo = (c_void_p * n)(*[o.ctypes.data for o in outs])
        TT = (c_long * n)(*[x.shape[0] for x in xs])
        self._ck(self._L.btkb200_chain_zelinski_batch(self._h, pp, TT, n, alpha, pf_type, min_frames, oo))

    # -- argument checks of the raw-pointer entry points: the C side trusts (pointer, T) pairs, so a wrong shape, dtype or
    #    a non-contiguous view would make it read or write past the end of the numpy buffer
    def _check_pcm_list(self, xs, dtype=np.float32, inner=None):
        for i, x in enumerate(xs):
            want_nd = 2 if inner is None else 3
            if (not isinstance(x, np.ndarray) or x.dtype != dtype or x.ndim != want_nd or x.shape[1] != self.C
                    or (inner is not None and x.shape[2] != inner) or not x.flags.c_contiguous):
                raise BtkError(EINVAL, f"recording {i}: need a C-contiguous {np.dtype(dtype).name} array of shape [T][{self.C}]"
                                       + (f"[{inner}]" if inner else "") + f", got {getattr(x, 'dtype', type(x))} {getattr(x, 'shape', '')}")

    def _check_out_list(self, Ts, outs):
        if len(outs) != len(Ts):
            raise BtkError(EINVAL, f"{len(Ts)} recordings but {len(outs)} output buffers")
        for i, (T, o) in enumerate(zip(Ts, outs)):
            need = self.chain_frames(int(T)) * self.D
            if (not isinstance(o, np.ndarray) or o.dtype != np.float32 or not o.flags.c_contiguous or o.size < need
                    or not o.flags.writeable):
                raise BtkError(EINVAL, f"output {i}: need a writable C-contiguous float32 array of at least {need} elements")

    def _check_snap(self, s):
        if s.ndim != 3 or s.shape[1:] != (self.B, self.C):
            raise BtkError(EINVAL, f"snapshots must be [F][{self.B}][{self.C}], got {s.shape}")

    # -- fused path (host numpy buffers)
    def chain(self, pcm) -> np.ndarray:
        x = np.ascontiguousarray(pcm, dtype=np.float32)
        if x.ndim == 1:
            x = x[:, None]
        if x.shape[1] != self.C:
            raise BtkError(EINVAL, f"pcm has {x.shape[1]} channels, plan has {self.C}")
        out = np.empty(self.chain_frames(x.shape[0]) * self.D, dtype=np.float32)
        self._ck(self._L.btkb200_chain(self._h, _p(x), x.shape[0], _p(out)))
        return out

    def chain_batch(self, pcms) -> list:
        xs = [np.ascontiguousarray(x, dtype=np.float32) for x in pcms]
        outs = [np.empty(self.chain_frames(x.shape[0]) * self.D, dtype=np.float32) for x in xs]
        self.chain_batch_into(xs, outs)
        return outs

    def chain_batch_into(self, xs, outs):
        """xs / outs: lists of C-contiguous float32 arrays (outs preallocated); no allocation here."""
        n = len(xs)
        self._check_pcm_list(xs)
        self._check_out_list([x.shape[0] for x in xs], outs)
        pp = (c_void_p * n)(*[x.ctypes.data for x in xs])
        oo = (c_void_p * n)(*[o.ctypes.data for o in outs])
        TT = (c_long * n)(*[x.shape[0] for x in xs])
        self._ck(self._L.btkb200_chain_batch(self._h, pp, TT, n, oo))

    def mvdr_chain_batch_into(self, xs, outs, forget=0.99, last_frame=-1, conjugate=True, load_abs=0.0, load_rel=0.0,
                              dThreshold=1e-8):
        """Per-recording covariance -> loading -> MVDR solve -> fused chain, all on the device.  Returns the number of
        identity-fallback bins per recording."""
        n = len(xs)
        self._check_pcm_list(xs)
        self._check_out_list([x.shape[0] for x in xs], outs)
        pp = (c_void_p * n)(*[x.ctypes.data for x in xs])
        oo = (c_void_p * n)(*[o.ctypes.data for o in outs])
        TT = (c_long * n)(*[x.shape[0] for x in xs])
        cfg = MvdrAdapt(forget, last_frame, 1 if conjugate else 0, load_abs, load_rel, dThreshold)
        nfb = np.zeros(n, dtype=np.int32)
        self._ck(self._L.btkb200_mvdr_chain_batch(self._h, pp, TT, n, ctypes.byref(cfg), oo, _p(nfb)))
        return nfb

    def mvdr_chain_batch(self, pcms, **kw):
        xs = [np.ascontiguousarray(x, dtype=np.float32) for x in pcms]
        outs = [np.empty(self.chain_frames(x.shape[0]) * self.D, dtype=np.float32) for x in xs]
        nfb = self.mvdr_chain_batch_into(xs, outs, **kw)
        return outs, nfb

    def chain_batch_pcm_into(self, raws, fmt: int, Ts, outs):
        """raws: C-contiguous arrays of raw interleaved PCM -- float32 [T][C] (PCM_F32), int16 [T][C] (PCM_S16) or
        uint8 [T][C][3] big-endian 24-bit (PCM_S24BE); Ts: samples per channel; outs preallocated float32."""
        n = len(raws)
        if fmt not in (PCM_F32, PCM_S16, PCM_S24BE):
            raise BtkError(EINVAL, f"unknown PCM format {fmt}")
        self._check_pcm_list(raws, {PCM_F32: np.float32, PCM_S16: np.int16, PCM_S24BE: np.uint8}[fmt], 3 if fmt == PCM_S24BE else None)
        if len(Ts) != n or any(int(t) != x.shape[0] for t, x in zip(Ts, raws)):
            raise BtkError(EINVAL, "Ts must give the number of samples per channel of every raw buffer")
        self._check_out_list(Ts, outs)
        pp = (c_void_p * n)(*[x.ctypes.data for x in raws])
        oo = (c_void_p * n)(*[o.ctypes.data for o in outs])
        TT = (c_long * n)(*[int(t) for t in Ts])
        self._ck(self._L.btkb200_chain_batch_pcm(self._h, pp, fmt, TT, n, oo))

    def chain_batch_pcm(self, raws, fmt: int) -> list:
        want = {PCM_F32: np.float32, PCM_S16: np.int16, PCM_S24BE: np.uint8}[fmt]
        xs = [np.ascontiguousarray(x, dtype=want) for x in raws]
        for x in xs:
            if x.ndim < 2 or x.shape[1] != self.C or (fmt == PCM_S24BE and (x.ndim != 3 or x.shape[2] != 3)):
                raise BtkError(EINVAL, f"raw pcm of shape {x.shape} does not match {self.C} channels / format {fmt}")
        outs = [np.empty(self.chain_frames(x.shape[0]) * self.D, dtype=np.float32) for x in xs]
        self.chain_batch_pcm_into(xs, fmt, [x.shape[0] for x in xs], outs)
        return outs

    def convert_pcm(self, raw, fmt: int) -> np.ndarray:
        """The ingest conversion alone (through the device): raw samples -> float32, same order."""
        want = {PCM_F32: np.float32, PCM_S16: np.int16, PCM_S24BE: np.uint8}[fmt]
        x = np.ascontiguousarray(raw, dtype=want)
        n = x.size // 3 if fmt == PCM_S24BE else x.size
        out = np.empty(n, dtype=np.float32)
        self._ck(self._L.btkb200_convert_pcm(self._h, _p(x), fmt, n, _p(out)))
        return out.reshape(x.shape[:-1] if fmt == PCM_S24BE else x.shape)

    # -- device-resident variants (raw device pointers as ints, e.g. torch.Tensor.data_ptr())
    def chain_batch_dev(self, d_pcm: int, pcm_off, T, out_off, d_out: int, stream: int = 0):
        po = np.ascontiguousarray(pcm_off, dtype=np.int64)
        tt = np.ascontiguousarray(T, dtype=np.int64)
        oo = np.ascontiguousarray(out_off, dtype=np.int64)
        LL = POINTER(c_longlong)
        self._ck(self._L.btkb200_chain_batch_dev(self._h, d_pcm, po.ctypes.data_as(LL), tt.ctypes.data_as(LL),
                                                 oo.ctypes.data_as(LL), po.size, d_out, stream))

    def analysis_dev(self, d_pcm: int, T: int, d_snap: int, stream: int = 0):
        self._ck(self._L.btkb200_analysis_dev(self._h, d_pcm, T, d_snap, stream))

    def beamform_dev(self, d_snap: int, F: int, d_Y: int, stream: int = 0):
        self._ck(self._L.btkb200_beamform_dev(self._h, d_snap, F, d_Y, stream))

    def synthesis_dev(self, d_Y: int, F: int, d_out: int, stream: int = 0):
        self._ck(self._L.btkb200_synthesis_dev(self._h, d_Y, F, d_out, stream))

    def launch_count(self) -> int:
        return int(self._L.btkb200_launch_count(self._h))

    def tune(self, chain_ws: int | None = None, cluster: int | None = None):
        """Launch-geometry knobs of the fused chain (btkb200_plan_tune): A/B runs and tests only."""
        if chain_ws is not None:
            self._ck(self._L.btkb200_plan_tune(self._h, TUNE_CHAIN_WS, int(chain_ws)))
        if cluster is not None:
            self._ck(self._L.btkb200_plan_tune(self._h, TUNE_CLUSTER, int(cluster)))

    def tuning(self) -> dict:
        """What the last fused-chain launch used: {'chain_ws': 0|1, 'cluster': n} (-1 before any launch)."""
        return {"chain_ws": int(self._L.btkb200_plan_tuning(self._h, TUNE_CHAIN_WS)),
                "cluster": int(self._L.btkb200_plan_tuning(self._h, TUNE_CLUSTER))}

    def sync(self):
        self._ck(self._L.btkb200_sync(self._h))


def design_analysis_prototype(M: int, m: int, r: int, wp_factor: float = 1.0, tau: int = -1, tolerance: float = 1e-7,
                              device: int = 0):
    """AnalysisOversampledDFTDesign(M, m, r, wpFactor, tau).design(tolerance) on the device.  Returns (h, err[3])."""
    L = lib()
    h = np.zeros(M * m, dtype=np.float64)
    err = np.zeros(3, dtype=np.float64)
    rc = L.btkb200_design_analysis_prototype(M, m, r, wp_factor, tau, tolerance, device, _p(h), _p(err))
    if rc != OK:
        raise BtkError(rc, (L.btkb200_last_error(None) or b"").decode())
    return h, err


def design_synthesis_prototype(h, M: int, m: int, r: int, v: float = 1.0, wp_factor: float = 1.0, tau: int = -1,
                               tolerance: float = 1e-7, device: int = 0):
    """SynthesisOversampledDFTDesign(h, M, m, r, v, wpFactor, tau).design(tolerance) on the device.  Returns (g, err[3])."""
    L = lib()
    hh = np.ascontiguousarray(h, dtype=np.float64)
    if hh.size != M * m:
        raise BtkError(EINVAL, f"Prototype sizes do not match ({hh.size} vs. {M * m}).")
    g = np.zeros(M * m, dtype=np.float64)
    err = np.zeros(3, dtype=np.float64)
    rc = L.btkb200_design_synthesis_prototype(_p(hh), M, m, r, v, wp_factor, tau, tolerance, device, _p(g), _p(err))
    if rc != OK:
        raise BtkError(rc, (L.btkb200_last_error(None) or b"").decode())
    return g, err


def design_analysis_nyquist(M: int, m: int, r: int, wp_factor: float = 1.0, tau: int = -1, tolerance: float = 1e-7,
                            device: int = 0):
    """AnalysisNyquistMDesign(M, m, r, wpFactor, tau).design(tolerance) on the device.  Returns (h, path)."""
    L = lib()
    h = np.zeros(M * m, dtype=np.float64)
    path = c_int(0)
    rc = L.btkb200_design_analysis_nyquist(M, m, r, c_double(wp_factor), tau, c_double(tolerance), device, _p(h), ctypes.byref(path))
    if rc != OK:
        raise BtkError(rc, (L.btkb200_last_error(None) or b"").decode())
    return h, int(path.value)


def design_synthesis_nyquist(h, M: int, m: int, r: int, wp_factor: float = 1.0, tau: int = -1, tolerance: float = 1e-7,
                             device: int = 0):
    """SynthesisNyquistMDesign(h, M, m, r, wpFactor, tau).design(tolerance) on the device.  Returns (g, path)."""
    L = lib()
    hh = np.ascontiguousarray(h, dtype=np.float64)
    if hh.size != M * m:
        raise BtkError(EINVAL, f"Prototype sizes do not match ({hh.size} vs. {M * m}).")
    g = np.zeros(M * m, dtype=np.float64)
    path = c_int(0)
    rc = L.btkb200_design_synthesis_nyquist(_p(hh), M, m, r, c_double(wp_factor), tau, c_double(tolerance), device, _p(g),
                                            ctypes.byref(path))
    if rc != OK:
        raise BtkError(rc, (L.btkb200_last_error(None) or b"").decode())
    return g, int(path.value)


def calc_delays_polar(azimuth: float, elevation: float, micpos_mm) -> np.ndarray:
    """calcDelaysPolar2 (src/superdirectiveBeamformer.cc:118-137): far-field delays in seconds, micpos [C][3] in mm."""
    mp = np.ascontiguousarray(micpos_mm, dtype=np.float64)
    if mp.ndim != 2 or mp.shape[1] != 3:
        raise BtkError(EINVAL, f"micpos must be [C][3], got {mp.shape}")
    d = np.zeros(mp.shape[0], dtype=np.float64)
    rc = lib().btkb200_calc_delays_polar(azimuth, elevation, _p(mp), mp.shape[0], _p(d))
    if rc:
        raise BtkError(rc, "btkb200_calc_delays_polar")
    return d


def calc_all_delays(x: float, y: float, z: float, micpos_mm) -> np.ndarray:
    """calcAllDelays (beamformer.cc:1214-1231); the source position is ignored, as in the reference."""
    mp = np.ascontiguousarray(micpos_mm, dtype=np.float64)
    if mp.ndim != 2 or mp.shape[1] != 3:
        raise BtkError(EINVAL, f"micpos must be [C][3], got {mp.shape}")
    d = np.zeros(mp.shape[0], dtype=np.float64)
    rc = lib().btkb200_calc_all_delays(x, y, z, _p(mp), mp.shape[0], _p(d))
    if rc:
        raise BtkError(rc, "btkb200_calc_all_delays")
    return d


def device_count() -> int:
    return int(lib().btkb200_device_count())


def chain_batch_multi(plans, xs, outs):
    """Round-robin the recordings over several plans (one per device); no inter-GPU traffic."""
    L = lib()
    n, npl = len(xs), len(plans)
    hp = (c_void_p * npl)(*[p._h.value for p in plans])
    pp = (c_void_p * n)(*[x.ctypes.data for x in xs])
    oo = (c_void_p * n)(*[o.ctypes.data for o in outs])
    TT = (c_long * n)(*[x.shape[0] for x in xs])
    rc = L.btkb200_chain_batch_multi(hp, npl, pp, TT, n, oo)
    if rc != OK:
        raise BtkError(rc, "chain_batch_multi failed: " + "; ".join((L.btkb200_last_error(p._h) or b"").decode() for p in plans))
