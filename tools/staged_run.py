#!/usr/bin/env python
"""Runs every kernel of the path once per configuration (for an ncu launch list):
analysis -> covariance (estimate) -> solve -> beamform -> synthesis, and the fused chain.
usage: python tools/staged_run.py [cfg2|cfg3|cfg4 ...]   (prints algorithmic bytes per kernel)"""
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
    for rep in range(2):     # second pass = warm
        plan.estimate_covariance(pcm, 0.99)                # analysis + covariance kernels
        plan.diag_load(1e-2 * float(np.real(np.trace(plan.get_covariance(40)))) / C)
        plan.solve_mvdr()                                  # solve kernel
        snap = plan.analysis(pcm)
        Y = plan.beamform(snap)                            # beamform kernel
        out = plan.synthesis(Y)                            # synthesis kernel
        out2 = plan.chain(pcm)                             # fused chain kernel
    F, B, D = snap.shape[0], plan.B, plan.D
    nblk = plan.nblk(T)
    print(f"{name}: C={C} M={M} T={T} F={F} | bytes: analysis {4*C*T + 8*C*F*B:.4g}  covariance {8*C*F*B + 16*B*C*C:.4g}  "
          f"beamform {8*C*F*B + 8*F*B:.4g}  synthesis {8*F*B + 4*nblk*D:.4g}  chain {4*C*T + 4*nblk*D:.4g}  solve {16*B*C*C + 32*B*C:.4g}")
    plan.close()
