"""Regenerates tests/golden/zelinski_*.npz from the COMPILED REFERENCE (oracle/_ref/libbtk_ref.so, which now includes the
reference's own postfilter/postfilter.cc): SubbandDS -> ZelinskiPostFilter -> OverSampledDFTSynthesisBank wired like
src/beamformerDS.cc:150-190.  Run HERE (needs /root/reference):   make -C oracle && python tests/golden/make_golden_zelinski.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import btk_oracle as bo  # noqa: E402


def main():
    P = np.load(os.path.join(HERE, "prototypes.npz"))
    ref = bo.CompiledReference()
    # name, M, m, r, dct, C, T, alpha, type, min_frames
    cases = [("abs_256_4_1_c4", 256, 4, 1, 0, 4, 2600, 0.6, 2, 0),
             ("real_512_2_2_c3", 512, 2, 2, 0, 3, 2000, 0.6, 1, 0),
             ("abs_256_4_1_c8_min5", 256, 4, 1, 0, 8, 2000, 0.9, 2, 5),
             ("nouse_512_2_3_c2", 512, 2, 3, 0, 2, 1200, 0.6, 0, 0)]
    for idx, (name, M, m, r, dct, C, T, alpha, typ, mf) in enumerate(cases):
        h, g = P[f"h_{M}_{m}_{r}"], P[f"g_{M}_{m}_{r}"]
        geo = bo.BankGeometry(M, m, r, dct)
        rng = np.random.default_rng(777 + idx)
        mp = np.stack([41.0 * np.arange(C), 7.0 * np.arange(C) ** 2, np.zeros(C)], axis=1).astype(np.float64)
        tau = bo.farfield_delays(mp, np.deg2rad(40.0), np.deg2rad(80.0))
        # a coherent source along the look direction plus sensor noise: the gains then spread over (floor, 1)
        t = np.arange(T) / 16000.0
        src = 3000.0 * np.sin(2 * np.pi * (300.0 + 2500.0 * t / t[-1]) * t)
        pcm = np.stack([np.interp(t - tau[c], t, src) for c in range(C)], axis=1)
        pcm = (pcm + 600.0 * rng.standard_normal((T, C))).astype(np.float32)
        res = ref.chain_zelinski(pcm, h, g, geo, tau, alpha, typ, mf)
        np.savez_compressed(os.path.join(HERE, f"zelinski_{name}.npz"), pcm=pcm, delays=tau,
                            geo=np.array([M, m, r, dct, C, T]), alpha=np.float64(alpha), pf_type=np.array(typ),
                            min_frames=np.array(mf), Ypf=res["Ypf"], Wpf=res["Wpf"][:, : geo.B], out=res["out"])
        print(name, "frames", res["frames"], "W range", res["Wpf"].min(), res["Wpf"].max())


if __name__ == "__main__":
    main()
