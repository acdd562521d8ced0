"""Regenerates tests/golden/*.npz from the reference itself.  Run HERE (needs /root/reference):

    make -C oracle && python tests/golden/make_golden.py

prototypes.npz : the h/g Nyquist(M) prototypes shipped as TEXT fixtures with the reference
                 (btk/examples/prototypes/Nyquist/M=*-m=*-r=*.m: h then g, 2N values, the format read by
                 getFilterCoeffs, btk/src/superdirectiveBeamformer.cc:24-47).  Input data, not code.
golden_*.npz   : outputs of the COMPILED REFERENCE (oracle/_ref/libbtk_ref.so = the reference's own .cc
                 files, see oracle/Makefile) on small seeded inputs.  These pin both the numpy oracle and
                 the CUDA path on machines where /root/reference does not exist (the GPU box).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import btk_oracle as bo  # noqa: E402

REF_PROTO = "/root/reference/btk/examples/prototypes/Nyquist"


def load_m(M, m, r):
    a = np.loadtxt(os.path.join(REF_PROTO, f"M={M}-m={m}-r={r}.m")).ravel()
    N = M * m
    assert a.size == 2 * N
    return a[:N].copy(), a[N:].copy()


def main():
    protos = {}
    for (M, m, r) in [(256, 4, 1), (512, 2, 2), (512, 2, 3)]:
        h, g = load_m(M, m, r)
        protos[f"h_{M}_{m}_{r}"] = h
        protos[f"g_{M}_{m}_{r}"] = g
    np.savez_compressed(os.path.join(HERE, "prototypes.npz"), **protos)

    ref = bo.CompiledReference()
    cases = []
    # (name, M, m, r, dct, C, T, mode, load)
    cases.append(("bank_256_4_1_dct0", 256, 4, 1, 0, 1, 1500, "ds", 0.0))
    cases.append(("bank_256_4_1_dct1", 256, 4, 1, 1, 1, 1000, "ds", 0.0))
    cases.append(("bank_512_2_2_dct2", 512, 2, 2, 2, 1, 1100, "ds", 0.0))
    cases.append(("bank_512_2_3_dct0", 512, 2, 3, 0, 1, 777, "ds", 0.0))
    cases.append(("ds_256_4_1_c4", 256, 4, 1, 0, 4, 1300, "ds", 0.0))
    cases.append(("ds_512_2_2_c3", 512, 2, 2, 0, 3, 900, "ds", 0.0))
    cases.append(("mvdr_512_2_2_c4", 512, 2, 2, 0, 4, 900, "mvdr", 0.1))
    cases.append(("mvdr_256_4_1_c6", 256, 4, 1, 0, 6, 700, "mvdr", 1.0))
    for idx, (name, M, m, r, dct, C, T, mode, load) in enumerate(cases):
        h = protos[f"h_{M}_{m}_{r}"]
        g = protos[f"g_{M}_{m}_{r}"]
        geo = bo.BankGeometry(M, m, r, dct)
        rng = np.random.default_rng(4242 + idx)
        pcm = (1000.0 * rng.standard_normal((T, C))).astype(np.float32)
        pcm += (3000.0 * np.sin(2 * np.pi * 440.0 * np.arange(T) / 16000.0)).astype(np.float32)[:, None]
        mp = np.stack([41.0 * np.arange(C), 7.0 * np.arange(C) ** 2, np.zeros(C)], axis=1).astype(np.float64)
        tau = bo.farfield_delays(mp, np.deg2rad(40.0), np.deg2rad(80.0))
        kw = dict(mode=mode)
        if mode == "mvdr":
            kw.update(micpos=mp, diag_load=load, inverse="double")
        res = ref.chain(pcm, h, g, geo, tau, **kw)
        out = dict(pcm=pcm, micpos=mp, delays=tau, geo=np.array([M, m, r, dct, C, T]), load=np.float64(load),
                   mode=np.array(1 if mode == "mvdr" else 0), X=res["X"], Y=res["Y"], out=res["out"], W=res["W"])
        if mode == "mvdr":
            resA = ref.chain(pcm, h, g, geo, tau, mode="mvdr", micpos=mp, diag_load=load, inverse="float",
                             want_snap=False)
            out["Y_floatsvd"] = resA["Y"]
            out["W_floatsvd"] = resA["W"]
        if C > 1:
            out["S_cpp"] = ref.spectral_matrix(pcm, h, geo, 0.95)
        np.savez_compressed(os.path.join(HERE, f"golden_{name}.npz"), **out)
        print(name, "frames", res["frames"], "out_frames", res["out_frames"])


if __name__ == "__main__":
    main()
