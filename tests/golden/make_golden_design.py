"""Regenerates tests/golden/design_*.npz from the COMPILED REFERENCE (oracle/_ref/libbtk_ref.so, which includes the reference's
own modulated/prototypeDesign.cc): AnalysisOversampledDFTDesign(M, m, r, wpFactor).design() and, from that h,
SynthesisOversampledDFTDesign(h, M, m, r, v, wpFactor).design() plus both calcError() vectors.
Run HERE (needs /root/reference):   make -C oracle && python tests/golden/make_golden_design.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import btk_oracle as bo  # noqa: E402


def main():
    ref = bo.CompiledReference()
    for (M, m, r, wp, v) in [(16, 2, 1, 1.0, 1.0), (64, 2, 1, 1.0, 1.0), (32, 4, 2, 1.0, 0.5), (128, 2, 2, 2.0, 1.0)]:
        h, g, eh, eg = ref.design_dehaan(M, m, r, wp, v, 1e-7)
        np.savez_compressed(os.path.join(HERE, f"design_{M}_{m}_{r}.npz"), geo=np.array([M, m, r]), wp=np.float64(wp),
                            v=np.float64(v), h=h, g=g, err_h=eh, err_g=eg)
        print((M, m, r), "eps", eh, eg)


if __name__ == "__main__":
    main()
