#!/usr/bin/env python
"""One call of the batched adaptive MVDR chain (btkb200_mvdr_chain_batch) at a bench.py geometry, for a launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/mvdr_adapt_run.py cfg4 64

prints the end-to-end time of the plain chain and of the adaptive one (host buffers) when run without ncu."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import btk_b200  # noqa: E402


def main():
    import torch
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    cfg = dict(bench.WORKLOADS[name])
    M, m, r, C = cfg["M"], cfg["m"], cfg["r"], cfg["C"]
    T = int(round(cfg["seconds"] * bench.FS))
    h, g = bench.prototypes(M, m, r)
    _, tau = bench.geometry(cfg)
    plan = btk_b200.Plan(M, m, r, C, h, g, device=0)
    plan.set_ds_weights(bench.FS, tau)
    x = bench.make_recording(cfg, tau, 0)
    hin = torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    xs = [hin.numpy()] * nb
    hout = torch.empty((nb, plan.chain_frames(T) * plan.D), dtype=torch.float32).pin_memory()
    outs = [hout[i].numpy() for i in range(nb)]
    kw = dict(forget=0.99, last_frame=int(bench.FS // plan.D), conjugate=True, load_abs=0.0, load_rel=1e-2)
    for _ in range(reps):
        t0 = time.perf_counter(); plan.chain_batch_into(xs, outs); torch.cuda.synchronize(); t1 = time.perf_counter()
        nfb = plan.mvdr_chain_batch_into(xs, outs, **kw); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"{name} x {nb}: plain chain {1e3 * (t1 - t0):.2f} ms, adaptive MVDR {1e3 * (t2 - t1):.2f} ms, fallback bins {int(nfb.sum())}", flush=True)


if __name__ == "__main__":
    main()
