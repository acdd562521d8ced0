#!/usr/bin/env python
"""BASELINE configs[4]: throughput sweep of the fused chain, device-resident, one GPU:
channels x subbands (64 utterances of T = 160 000 samples) and a batch sweep for two shapes.
Prints a markdown table (kernel time from CUDA events on the launching stream, algorithmic bytes of SURVEY 8d against the
measured HBM peak).   usage (GPU box): python tools/sweep_bench.py > gpurun_out/sweep.md"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import btk_b200
wl = btk_b200.workloads
P = np.load(os.path.join(ROOT, "tests", "golden", "prototypes.npz"))
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
PEAK = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
T, FS = 160000, 16000.0
dev = torch.device("cuda", 0)


def run(M, m, r, C, nb, steps=10):
    key = f"h_{M}_{m}_{r}"
    h, g = (P[key], P[f"g_{M}_{m}_{r}"]) if key in P.files else wl.designed_prototype(M, m, r)   # no fixture: designed on the device (de Haan)
    plan = btk_b200.Plan(M, m, r, C, h, g)
    tau = wl.farfield_delays(wl.linear_array(C, 20.0), np.deg2rad(30), np.deg2rad(90))
    plan.set_ds_weights(FS, tau)
    nblk, D = plan.chain_frames(T), plan.D
    x = torch.from_numpy(wl.noise_recording(T, C, seed=C * 7 + M, sigma=800.0).reshape(-1)).to(dev)
    d_in = x.repeat(nb, 1).contiguous()
    d_out = torch.zeros((nb, nblk * D), dtype=torch.float32, device=dev)
    po = np.arange(nb, dtype=np.int64) * (T * C); oo = np.arange(nb, dtype=np.int64) * (nblk * D); Ts = np.full(nb, T, np.int64)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        plan.chain_batch_dev(d_in.data_ptr(), po, Ts, oo, d_out.data_ptr(), st)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record(); plan.chain_batch_dev(d_in.data_ptr(), po, Ts, oo, d_out.data_ptr(), st); b.record()
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    plan.close()
    bytes_ = nb * (4.0 * C * T + 4.0 * nblk * D)
    return ms, nb * C * T / FS / (ms * 1e-3), bytes_ / (ms * 1e-3) / 1e9 / PEAK


print("| M (m, r) | channels | utterances | kernel ms | M channel-s/s | frac of HBM roof |\n|---|---|---|---|---|---|")
for (M, m, r) in [(128, 2, 1), (256, 4, 1), (512, 2, 2), (1024, 2, 1)]:
    for C in (4, 8, 16, 32, 64):
        nb = 64
        ms, rate, frac = run(M, m, r, C, nb)
        print(f"| {M} ({m}, {r}) | {C} | {nb} | {ms:.3f} | {rate / 1e6:.2f} | {frac:.3f} |", flush=True)
print("\n| shape | utterances | kernel ms | M channel-s/s | frac |\n|---|---|---|---|---|")
for (M, m, r, C) in [(256, 4, 1, 8), (512, 2, 2, 64)]:
    for nb in (1, 8, 64, 512 if C == 8 else 128):
        ms, rate, frac = run(M, m, r, C, nb)
        print(f"| M={M} C={C} | {nb} | {ms:.3f} | {rate / 1e6:.2f} | {frac:.3f} |", flush=True)
