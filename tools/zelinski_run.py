#!/usr/bin/env python
"""Runs the post-filter chain once per configuration (for an ncu launch list): analysis -> beamform + Zelinski statistics ->
scan -> synthesis.   usage: python tools/zelinski_run.py [cfg2|cfg3|cfg4 ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import btk_b200
wl = btk_b200.workloads
CFG = {"cfg2": (256, 4, 1, 8, 60.0, "circ"), "cfg3": (512, 2, 2, 16, 10.0, "lin41"), "cfg4": (512, 2, 2, 64, 10.0, "lin20")}
P = np.load(os.path.join(ROOT, "tests", "golden", "prototypes.npz"))
for name in (sys.argv[1:] or ["cfg2", "cfg3", "cfg4"]):
    M, m, r, C, secs, geom = CFG[name]
    T = int(secs * 16000)
    h, g = P[f"h_{M}_{m}_{r}"], P[f"g_{M}_{m}_{r}"]
    mp = wl.circular_array(C) if geom == "circ" else wl.linear_array(C, 41.0 if geom == "lin41" else 20.0)
    tau = wl.farfield_delays(mp, np.deg2rad(30), np.deg2rad(90))
    pcm = wl.noise_recording(T, C, seed=1, sigma=300.0)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    plan.set_ds_weights(16000.0, tau)
    for rep in range(2):
        out = plan.chain_zelinski(pcm, 0.6, 2, 0)
    print(name, out.shape)
    plan.close()
