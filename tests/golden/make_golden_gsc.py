"""Regenerates tests/golden/gsc_*.npz from the COMPILED REFERENCE (oracle/_ref/libbtk_ref.so): SubbandGSC with
calcGSCWeights + setActiveWeights_f for every bin (beamformer/beamformer.cc:1296-1447) -> synthesis.
Run HERE (needs /root/reference):   make -C oracle && python tests/golden/make_golden_gsc.py"""
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
    # name, M, m, r, dct, C, T, normalize
    cases = [("256_4_1_c4", 256, 4, 1, 0, 4, 1500, False), ("512_2_2_c6_norm", 512, 2, 2, 0, 6, 1200, True),
             ("256_4_1_c2", 256, 4, 1, 1, 2, 900, False)]
    for idx, (name, M, m, r, dct, C, T, norm) in enumerate(cases):
        h, g = P[f"h_{M}_{m}_{r}"], P[f"g_{M}_{m}_{r}"]
        geo = bo.BankGeometry(M, m, r, dct)
        rng = np.random.default_rng(999 + idx)
        mp = np.stack([41.0 * np.arange(C), 7.0 * np.arange(C) ** 2, np.zeros(C)], axis=1).astype(np.float64)
        tau = bo.farfield_delays(mp, np.deg2rad(40.0), np.deg2rad(80.0))
        pcm = (1000.0 * rng.standard_normal((T, C))).astype(np.float32)
        wa = 0.08 * (rng.standard_normal((geo.B, C - 1)) + 1j * rng.standard_normal((geo.B, C - 1)))
        res = ref.chain_gsc(pcm, h, g, geo, tau, wa, norm)
        np.savez_compressed(os.path.join(HERE, f"gsc_{name}.npz"), pcm=pcm, delays=tau, wa=wa,
                            geo=np.array([M, m, r, dct, C, T]), normalize=np.array(int(norm)), Y=res["Y"],
                            out=res["out"], Bm=res["Bm"], wq=res["wq"])
        print(name, "frames", res["frames"])


if __name__ == "__main__":
    main()
