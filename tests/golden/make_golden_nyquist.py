#!/usr/bin/env python
"""Golden fixtures of the Nyquist(M)-constrained prototype designs: outputs of the COMPILED REFERENCE
(oracle/_ref: AnalysisNyquistMDesign / SynthesisNyquistMDesign of modulated/prototypeDesign.cc:955-1119, run on top of
oracle/gsl_shim) for the BASELINE geometries and a few small ones.

    python tests/golden/make_golden_nyquist.py [M m r tol] ...      (default: the list below; needs /root/reference)

The L = 1024 designs take minutes each on one core (the stand-in SVD is a serial Jacobi iteration)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import btk_oracle as bo  # noqa: E402

DEFAULT = [(16, 2, 1, 1e-7), (32, 4, 1, 1e-7), (64, 2, 2, 1e-7), (64, 2, 1, 0.2), (128, 2, 1, 1e-7), (256, 4, 1, 1e-7), (512, 2, 2, 1e-7)]


def main():
    a = sys.argv[1:]
    cases = [(int(a[i]), int(a[i + 1]), int(a[i + 2]), float(a[i + 3])) for i in range(0, len(a), 4)] if a else DEFAULT
    ref = bo.CompiledReference()
    for M, m, r, tol in cases:
        t = time.time()
        h, g = ref.design_nyquist(M, m, r, 1.0, tol)
        name = os.path.join(HERE, f"design_nyquist_{M}_{m}_{r}_{tol:g}.npz")
        np.savez_compressed(name, h=h, g=g, geo=np.array([M, m, r]), tol=tol, wp_factor=1.0)
        print(f"{name}: {time.time() - t:.1f} s", flush=True)


if __name__ == "__main__":
    main()
