#!/usr/bin/env python
"""Device-resident timing of the fused chain only (no end-to-end legs, no CPU arm): the quick A/B loop while tuning.

    python tools/kbench.py cfg2 cfg3 cfg4:ws=1,cluster=4 cfg4:ws=0 C=64,M=1024,m=2,r=1,batch=16:cluster=8 ...

A spec is a bench.py workload name or a comma list of overrides (C, M, m, r, batch, seconds), optionally followed by
`:` and plan tuning knobs (ws = 0 | 1, cluster = n).  Prints one line per spec: kernel ms (mean of --steps launches, CUDA
events on the launching stream), algorithmic GB/s and the fraction of the measured HBM peak."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def parse(spec):
    wl, _, tune = spec.partition(":")
    if wl in bench.WORKLOADS:
        cfg = dict(bench.WORKLOADS[wl])
    else:
        cfg = dict(M=512, m=2, r=2, C=16, seconds=10.0, batch=32, geom="linear20", desc=wl)
        for kv in wl.split(","):
            k, v = kv.split("=")
            cfg[k] = float(v) if k == "seconds" else int(v)
    knobs = {}
    if tune:
        for kv in tune.split(","):
            k, v = kv.split("=")
            knobs[k] = int(v)
    return cfg, knobs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("specs", nargs="+")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import torch

    import btk_b200

    peak = 6556.5
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peak = float(json.load(open(pp))["hbm_gbs"])
    dev = torch.device("cuda", 0)
    for spec in args.specs:
        cfg, knobs = parse(spec)
        M, m, r, C, nb = cfg["M"], cfg["m"], cfg["r"], cfg["C"], cfg["batch"]
        T = int(round(cfg["seconds"] * bench.FS))
        h, g = bench.prototypes(M, m, r)
        _, tau = bench.geometry(cfg)
        plan = btk_b200.Plan(M, m, r, C, h, g, device=0)
        plan.set_ds_weights(bench.FS, tau)
        try:
            if "ws" in knobs:
                plan.tune(chain_ws=knobs["ws"])
            if "cluster" in knobs:
                plan.tune(cluster=knobs["cluster"])
        except btk_b200.BtkError as e:
            print(f"{spec:48s} unsupported: {e.msg}")
            plan.close()
            continue
        nblk, D = plan.nblk(T), plan.D
        n_in, n_out = T * C, nblk * D
        base = torch.from_numpy(bench.make_recording(cfg, tau, 0).reshape(-1)).to(dev)
        d_in = base.repeat(nb).contiguous()
        d_out = torch.zeros(nb * n_out, dtype=torch.float32, device=dev)
        pcm_off = np.arange(nb, dtype=np.int64) * n_in
        out_off = np.arange(nb, dtype=np.int64) * n_out
        Ts = np.full(nb, T, dtype=np.int64)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            plan.chain_batch_dev(d_in.data_ptr(), pcm_off, Ts, out_off, d_out.data_ptr(), st)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in evs:
            a.record()
            plan.chain_batch_dev(d_in.data_ptr(), pcm_off, Ts, out_off, d_out.data_ptr(), st)
            b.record()
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        alg = nb * (4.0 * C * T + 4.0 * nblk * D)
        gbs = alg / (ms * 1e-3) / 1e9
        tu = plan.tuning()
        print(f"{spec:48s} ws={tu['chain_ws']} cluster={tu['cluster']}  {ms:8.4f} ms  {gbs:8.1f} GB/s  frac {gbs / peak:.4f}  "
              f"{nb * C * cfg['seconds'] / (ms * 1e-3) / 1e6:7.2f} M ch-s/s", flush=True)
        plan.close()
        del d_in, d_out


if __name__ == "__main__":
    main()
