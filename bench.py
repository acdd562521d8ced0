#!/usr/bin/env python
"""bench.py -- throughput of the hot path (analysis -> SubbandDS -> synthesis) in channel-audio-seconds/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the fused chain over this rank's batch of synthetic array recordings
(BASELINE.json configs[1] by default: 8-mic circular array, SubbandDS, M=256 m=4 r=1, 16 kHz, 60 s).
Recordings are independent, so ranks share nothing on the data path (weak scaling: fixed batch per GPU);
torch.distributed is only used for the barrier and the max-over-ranks time.

  value      device-resident: inputs already in HBM, CUDA events around the K timed steps on the launching stream; the K
             steps are ONE CUDA graph of K launches of btkb200_chain_batch_dev (captured once, one untimed replay, then
             the timed replay), so that a slow host is not measured in place of 0.25-ms kernels; kernel_ms = total / K;
             --no-graph times K direct launches with an event pair each ("timed_as" in the line says which)
  e2e        through the host-buffer C-ABI call (btkb200_chain_batch): pinned host buffers, H2D of every
             input and D2H of every output inside the timed region
  roofline   algorithmic bytes of the fused chain (4 C T + 4 nblk D per recording) / launch time,
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU chain (oracle/_ref, compiled from /root/reference) on the host cores
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

FS = 16000.0
METRIC = "channel-audio-seconds/sec (analysis+beamform+synthesis)"
UNIT = "channel-s/s"

WORKLOADS = {
    # name: (M, m, r, C, seconds, utterances per GPU, geometry, description)
    "cfg2": dict(M=256, m=4, r=1, C=8, seconds=60.0, batch=16, geom="circular",
                 desc="BASELINE configs[1]: 8-mic circular array SubbandDS, M=256 m=4 r=1, 16 kHz, 60 s source at known DOA"),
    "cfg3": dict(M=512, m=2, r=2, C=16, seconds=10.0, batch=64, geom="linear41",
                 desc="BASELINE configs[2] geometry: 16-mic linear array, fixed-weight MVDR/DS apply, M=512 m=2 r=2, 16 kHz, 10 s"),
    "m1024": dict(M=1024, m=2, r=1, C=8, seconds=10.0, batch=64, geom="circular",
                  desc="BASELINE configs[4] sweep point: 8 channels, M=1024 m=2 r=1 (Kaiser prototype), 16 kHz, 10 s utterances"),
    "m128": dict(M=128, m=2, r=1, C=8, seconds=10.0, batch=64, geom="circular",
                 desc="BASELINE configs[4] sweep point: 8 channels, M=128 m=2 r=1 (Kaiser prototype), 16 kHz, 10 s utterances"),
    "cfg4": dict(M=512, m=2, r=2, C=64, seconds=10.0, batch=32, geom="linear20",
                 desc="BASELINE configs[3] geometry: 64-ch Mark-III-style array, M=512 m=2 r=2, 16 kHz, 10 s utterances"),
}


def prototypes(M, m, r):
    import btk_b200

    P = np.load(os.path.join(ROOT, "tests", "golden", "prototypes.npz"))
    key = f"h_{M}_{m}_{r}"
    if key in P.files:
        return P[key], P[f"g_{M}_{m}_{r}"]
    # no shipped fixture for this geometry (M = 128, 1024, ...): design the pair on the device with the reference's own
    # procedure; the CPU reference arm has no device and keeps the Kaiser stand-in (it is only ever run on
    # the fixture geometries)
    try:
        return btk_b200.workloads.designed_prototype(M, m, r)
    except btk_b200.BtkError:
        return btk_b200.workloads.kaiser_prototype(M, m, r)


def geometry(wl_cfg):
    import btk_b200

    wl = btk_b200.workloads
    C = wl_cfg["C"]
    if wl_cfg["geom"] == "circular":
        mp_ = wl.circular_array(C, 100.0)
        tau = wl.farfield_delays(mp_, np.deg2rad(60.0), np.deg2rad(90.0))
    else:
        mp_ = wl.linear_array(C, 41.0 if wl_cfg["geom"] == "linear41" else 20.0)
        tau = wl.farfield_delays(mp_, np.deg2rad(30.0), np.deg2rad(90.0))
    return mp_, tau


def make_recording(wl_cfg, tau, index, config_id=2):
    import btk_b200

    T = int(round(wl_cfg["seconds"] * FS))
    return btk_b200.workloads.array_recording(T, tau, seed=20240 + 1000 * config_id + index)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md clocks line)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    """One single-threaded process per host core (the reference has no threading and global pools)."""
    wl_cfg, indices, use_ref = args
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)   # the reference's constructors chat on stdout
    os.dup2(devnull, 2)
    import btk_oracle as bo

    M, m, r, C = wl_cfg["M"], wl_cfg["m"], wl_cfg["r"], wl_cfg["C"]
    h, g = prototypes(M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    _, tau = geometry(wl_cfg)
    ref = bo.CompiledReference() if use_ref else None
    W = None if use_ref else bo.ds_weights(tau, FS, M)
    pcms = [make_recording(wl_cfg, tau, i) for i in indices]
    t0 = time.perf_counter()
    for pcm in pcms:
        if use_ref:
            ref.chain(pcm, h, g, geo, tau, want_snap=False, want_Y=False)
        else:
            bo.chain(pcm, h, g, geo, W)
    return time.perf_counter() - t0, len(pcms)


def cpu_chain_rate(wl_cfg, seconds_per_rec, recs_per_core, cores, pool):
    """channel-audio-seconds per second of the reference CPU chain using `cores` processes."""
    import btk_oracle as bo

    use_ref = bo.CompiledReference.available()
    sample_cfg = dict(wl_cfg, seconds=seconds_per_rec)
    jobs = [(sample_cfg, list(range(k * recs_per_core, (k + 1) * recs_per_core)), use_ref) for k in range(cores)]
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)          # generation of the inputs is outside each worker's timer
    units = sum(r[1] for r in res) * wl_cfg["C"] * seconds_per_rec
    return units / busy, ("reference" if use_ref else "port"), wall


def run_reference_arm(args, wl_cfg, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    secs = wl_cfg["seconds"]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_chain_rate(wl_cfg, 1.0, 1, cores, pool)
        rates, kind = [], "port"
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rate, kind, _ = cpu_chain_rate(wl_cfg, secs, 1, cores, pool)
            rates.append(rate)
        total = time.perf_counter() - t0
    value = float(np.mean(rates))
    sample = f"{cores} processes x 1 recording of {wl_cfg['C']} ch x {secs:g} s per step ({kind}: " + (
        "oracle/_ref = reference .cc files compiled -O2, GSL-radix-2 branch" if kind == "reference" else "numpy restatement") + ")"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl_cfg, args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(wl_cfg, args):
    return {"workload": wl_cfg["desc"], "channels": wl_cfg["C"], "M": wl_cfg["M"], "m": wl_cfg["m"], "r": wl_cfg["r"],
            "fs": FS, "seconds_per_utterance": wl_cfg["seconds"], "utterances_per_gpu": wl_cfg["batch"],
            "beamformer": "SubbandDS (delay-and-sum at the true DOA)", "parallelism": f"{args.gpus} independent shards, no collective",
            "l2": "per-step input+output exceeds the 126 MB L2 (no flush needed)"}


def bind_to_gpu_numa_node(index):
    """Pin this rank (and therefore the first touch of its pinned staging buffers) to the host cores NVML reports as
    local to its GPU: with one rank per GPU the H2D/D2H copies of the end-to-end path then stay on the local socket.
    Returns the previous affinity (restored before the CPU baseline runs)."""
    try:
        import pynvml

        old = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (max(old) + 64) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(hnd, words)
        cpus = {64 * w + b for w, v in enumerate(mask) for b in range(64) if (int(v) >> b) & 1} & old
        if cpus:
            os.sched_setaffinity(0, cpus)
        return old
    except Exception:
        return None


# --------------------------------------------------------------------------------------------- our arm
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # 148 SMs x 128 FP32 lanes x FMA at the 1965 MHz boost clock


def chain_flops_per_channel_sample(M, m, r, C):
    """analysis 2mR + 2.5 R log2 M, weight apply 4R, synthesis (2mR + 2.5 R log2 M) / C."""
    R = 1 << r
    bank = 2.0 * m * R + 2.5 * R * np.log2(M)
    return bank + 4.0 * R + bank / C


# the extra configurations the default run also measures (BASELINE.json configs[2], configs[3]); batch = utterances in total,
# split over the ranks (strong scaling) -- the primary line keeps its fixed batch per GPU (weak scaling)
EXTRA = [
    ("cfg3_mvdr", "cfg3", "mvdr", 64),
    ("cfg4_ds", "cfg4", "ds", 1024),
    ("cfg4_mvdr", "cfg4", "mvdr", 1024),
]
MVDR_LOAD = {16: 0.1, 64: 1.0}     # absolute diagonal loading of the diffuse-noise model (keeps the reference's
                                   # float-SVD inverse and the exact solve within 1e-6 of each other)


def _reference_output_worker(job):
    """One recording through the compiled reference chain (the parity check of the timed output)."""
    wl_cfg, mode, index, config_id = job
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    os.dup2(devnull, 2)
    import btk_oracle as bo

    M, m, r, C = wl_cfg["M"], wl_cfg["m"], wl_cfg["r"], wl_cfg["C"]
    h, g = prototypes(M, m, r)
    geo = bo.BankGeometry(M, m, r, 0)
    mp_, tau = geometry(wl_cfg)
    pcm = make_recording(wl_cfg, tau, index, config_id)
    if bo.CompiledReference.available():
        ref = bo.CompiledReference()
        if mode == "mvdr":
            res = ref.chain(pcm, h, g, geo, tau, mode="mvdr", micpos=mp_, diag_load=MVDR_LOAD.get(C, 0.1), inverse="double",
                            want_snap=False, want_Y=False)
        else:
            res = ref.chain(pcm, h, g, geo, tau, want_snap=False, want_Y=False)
        return res["out"], "reference"
    W = bo.ds_weights(tau, FS, M)
    if mode == "mvdr":
        W = bo.mvdr_weights(bo.diagonal_load(bo.diffuse_coherence(mp_, FS, M), MVDR_LOAD.get(C, 0.1)), W)
    return bo.chain(pcm, h, g, geo, W)[2], "port"


def snr_db(x, ref):
    x = np.asarray(x, np.float64); ref = np.asarray(ref, np.float64)
    n = min(x.size, ref.size)
    err = float(np.sum((x[:n] - ref[:n]) ** 2)) + 1e-300
    return 10.0 * np.log10(float(np.sum(ref[:n] ** 2)) / err)


def measure(wl_cfg, mode, nb, args, rank, world, local_rank, dev, config_id, steps, full_host_inputs, with_s16, adaptive,
            sampler=None):
    """Time one configuration on this rank.  Returns a dict of this rank's raw numbers (times in seconds / ms)."""
    import torch
    import torch.distributed as dist

    import btk_b200

    M, m, r, C = wl_cfg["M"], wl_cfg["m"], wl_cfg["r"], wl_cfg["C"]
    T = int(round(wl_cfg["seconds"] * FS))
    h, g = prototypes(M, m, r)
    mp_, tau = geometry(wl_cfg)
    plan = btk_b200.Plan(M, m, r, C, h, g, device=local_rank)
    plan.set_ds_weights(FS, tau)
    if mode == "mvdr":
        # SubbandMVDR: diffuse-noise coherence + diagonal loading -> per-bin solve -> weights (beamformer.cc:2392-2581)
        plan.set_diffuse_noise_model(mp_, FS)
        plan.diag_load(MVDR_LOAD.get(C, 0.1))
        if plan.solve_mvdr(FS, 1e-8) != 0:
            raise SystemExit("bench.py: MVDR solve fell back to identity weights")
    nblk, D = plan.nblk(T), plan.D
    n_in, n_out = T * C, nblk * D

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # this rank's shard of the job's world * nb utterances (round-robin ownership, btk_b200.sharding -- the same rule
    # btkb200_chain_batch_multi applies inside one process); a few distinct signals of it, reused round-robin
    distinct = min(nb, 4)
    mine = btk_b200.sharding.shard(world * nb, rank, world)
    base = [make_recording(wl_cfg, tau, mine[i], config_id) for i in range(distinct)]
    n_host = nb if full_host_inputs else distinct
    h_in = torch.empty((n_host, n_in), dtype=torch.float32, pin_memory=True)
    for i in range(n_host):
        h_in[i].copy_(torch.from_numpy(base[i % distinct].reshape(-1)))
    h_out = torch.empty((nb, n_out), dtype=torch.float32, pin_memory=True)
    d_in = torch.empty((nb, n_in), dtype=torch.float32, device=dev)
    for i in range(nb):
        d_in[i].copy_(h_in[i % n_host], non_blocking=True)
    d_out = torch.zeros((nb, n_out), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    pcm_off = np.arange(nb, dtype=np.int64) * n_in
    out_off = np.arange(nb, dtype=np.int64) * n_out
    Ts = np.full(nb, T, dtype=np.int64)
    stream = torch.cuda.current_stream().cuda_stream

    def step_dev():
        plan.chain_batch_dev(d_in.data_ptr(), pcm_off, Ts, out_off, d_out.data_ptr(), stream)

    # ---- device-resident timing.  The clock sampler (primary configuration only) runs from before the warm-up to after
    # the timed steps; the untimed warm-up is stretched to ~0.8 s of the same launches so that the samples show the clocks
    # this workload sustains going into the timed steps.
    if sampler is not None:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_dev()
    torch.cuda.synchronize()
    if sampler is not None:
        t_load = time.perf_counter()
        while time.perf_counter() - t_load < 0.8:
            for _ in range(50):
                step_dev()
            torch.cuda.synchronize()
    # The K timed steps are ONE CUDA graph of K launches of the C-ABI call (captured once, replayed inside the timed
    # region): a step is 0.25 ms, and a host that is slowed down for a few milliseconds (one run in five on a shared box
    # showed 0.34 ms per step with a 0.26-ms kernel: the queue ran dry) would otherwise be measured instead of the GPU.
    # The kernel's average launch duration is then total / K -- an upper bound of the per-launch figure (it includes the
    # gaps between the nodes).  --no-graph (or a failed capture) times the K direct calls with an event pair each.
    graph = None
    if not args.no_graph:
        try:
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                cs = torch.cuda.current_stream().cuda_stream
                for _ in range(steps):
                    plan.chain_batch_dev(d_in.data_ptr(), pcm_off, Ts, out_off, d_out.data_ptr(), cs)
            torch.cuda.synchronize()
            g_.replay()                      # one untimed replay: instantiation / upload costs stay out of the timed one
            torch.cuda.synchronize()
            graph = g_
        except Exception as exc:             # capture refused: direct launches below
            sys.stderr.write(f"bench.py: CUDA graph capture failed ({exc}); timing direct launches\n")
            graph = None
            torch.cuda.synchronize()
    barrier()
    l0 = plan.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record()
    if graph is not None:
        graph.replay()
    else:
        for a, b in evs:
            a.record()
            step_dev()
            b.record()
    t_end.record()
    barrier()
    torch.cuda.synchronize()
    total_ms = t_begin.elapsed_time(t_end)
    res = {"launches": steps if graph is not None else plan.launch_count() - l0, "total_ms": total_ms,
           "kern_ms": total_ms / steps if graph is not None else float(np.mean([a.elapsed_time(b) for a, b in evs])),
           "timed_as": "one CUDA graph of K launches" if graph is not None else "K direct launches", "tuning": plan.tuning()}
    if sampler is not None:
        res["clocks"] = sampler.stop()

    # ---- end to end through the host-buffer C-ABI call (H2D + kernels + D2H inside)
    xs = [h_in[i % n_host].numpy().reshape(T, C) for i in range(nb)]
    outs = [h_out[i].numpy() for i in range(nb)]
    e2e_steps = max(2, min(steps, 5))
    plan.chain_batch_into(xs, outs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.chain_batch_into(xs, outs)
    torch.cuda.synchronize()
    res["e2e_s"] = (time.perf_counter() - t0) / e2e_steps
    checksum = float(h_out[0, : 4 * D].double().abs().sum())
    if not np.isfinite(checksum) or checksum == 0.0:
        raise SystemExit("bench.py: end-to-end output is empty or not finite")
    res["same"] = bool(torch.equal(d_out[0].cpu(), h_out[0]))
    res["out0"] = h_out[0].numpy().copy()          # recording `first` of this shard: parity-checked against the reference

    # ---- the same call fed with 16-bit PCM (what the recordings are on disk: the reference converts 16-bit WAV to
    # float on the host, feature/feature.cc:273, 868-896); the conversion runs on the device, H2D bytes halve
    res["e2e16_s"] = 0.0
    if with_s16:
        h_in16 = torch.empty((n_host, n_in), dtype=torch.int16, pin_memory=True)
        for i in range(n_host):
            h_in16[i].copy_(torch.from_numpy(np.round(base[i % distinct].reshape(-1)).clip(-32768, 32767).astype(np.int16)))
        raws = [h_in16[i % n_host].numpy().reshape(T, C) for i in range(nb)]
        plan.chain_batch_pcm_into(raws, btk_b200._capi.PCM_S16, [T] * nb, outs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.chain_batch_pcm_into(raws, btk_b200._capi.PCM_S16, [T] * nb, outs)
        torch.cuda.synchronize()
        res["e2e16_s"] = (time.perf_counter() - t0) / e2e_steps
        if not np.isfinite(float(h_out[0, : 4 * D].double().abs().sum())):
            raise SystemExit("bench.py: 16-bit end-to-end output is not finite")

    # ---- per-utterance adaptive MVDR, end to end (covariance on a 1-s lead-in -> loading -> solve -> chain)
    res["adaptive_s"] = 0.0
    if adaptive:
        lead = int(FS // D)          # frames of the first second
        kw = dict(forget=0.99, last_frame=lead, conjugate=True, load_abs=0.0, load_rel=1e-2)
        plan.mvdr_chain_batch_into(xs, outs, **kw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            nfb = plan.mvdr_chain_batch_into(xs, outs, **kw)
        torch.cuda.synchronize()
        res["adaptive_s"] = (time.perf_counter() - t0) / e2e_steps
        if int(nfb.sum()) != 0 or not np.isfinite(float(h_out[0, : 4 * D].double().abs().sum())):
            raise SystemExit("bench.py: adaptive MVDR produced fallback bins or non-finite output")

    # ---- optional: the shipped drivers' chain with the Zelinski post-filter, end to end
    res["pf_s"] = 0.0
    if args.postfilter and mode == "ds":
        plan.chain_zelinski_batch_into(xs, outs, 0.6, 2, 0)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.chain_zelinski_batch_into(xs, outs, 0.6, 2, 0)
        torch.cuda.synchronize()
        res["pf_s"] = (time.perf_counter() - t0) / e2e_steps
    res.update(nblk=nblk, D=D, T=T, n_in=n_in, n_out=n_out, nb=nb, first=mine[0])
    plan.close()
    del d_in, d_out, h_in, h_out
    torch.cuda.empty_cache()
    return res


def summarise(wl_cfg, mode, res, world, steps, peak, peak_src, traffic):
    """This configuration's numbers, after the max over ranks: whole-job throughput, both rooflines, end to end."""
    M, m, r, C = wl_cfg["M"], wl_cfg["m"], wl_cfg["r"], wl_cfg["C"]
    nb, T, nblk, D = res["nb"], res["T"], res["nblk"], res["D"]
    units = nb * C * wl_cfg["seconds"]                        # channel-audio-seconds one rank processes per step
    alg_bytes = nb * (4.0 * C * T + 4.0 * nblk * D)           # SURVEY 8d fused-chain bytes per launch
    flops = nb * C * T * chain_flops_per_channel_sample(M, m, r, C)
    achieved = alg_bytes / (res["kern_ms"] * 1e-3) / 1e9
    tf = flops / (res["kern_ms"] * 1e-3) / 1e12
    ms_per_step = res["total_ms"] / steps
    out = {
        "value": world * units / (ms_per_step * 1e-3), "unit": UNIT, "ms_per_step": ms_per_step,
        "beamformer": "SubbandMVDR (diffuse-noise model + diagonal loading, per-bin solve on the device, then fixed weights)"
                      if mode == "mvdr" else "SubbandDS (delay-and-sum at the true DOA)",
        "utterances": world * nb, "channels": C, "M": M, "m": m, "r": r,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel_ms": res["kern_ms"],
                     "kernel": ("btk_chain_ws_kernel" if res["tuning"]["chain_ws"] else "btk_chain_kernel") + f"<{M},{1 << r}>",
                     "cluster": res["tuning"]["cluster"]},
        "roofline_fp32": {"bound": "fp32", "achieved": tf, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tf / FP32_PEAK_TFLOPS,
                          "flop_per_channel_sample": chain_flops_per_channel_sample(M, m, r, C),
                          "peak_source": "148 SMs x 128 lanes x 2 flop x 1.965 GHz (SURVEY 8d)"},
        "e2e": {"value": world * units / res["e2e_s"], "unit": UNIT, "h2d_bytes_per_step": int(nb * res["n_in"] * 4),
                "d2h_bytes_per_step": int(nb * res["n_out"] * 4), "ms_per_step": res["e2e_s"] * 1e3,
                "per_gpu": units / res["e2e_s"], "matches_device_resident_output": res["same"],
                # what bounds this leg: bytes over the host link per rank and second, and the fraction of the one-GPU
                # end-to-end rate each rank keeps (one GPU moves 553 MB in 9.3 ms = 59 GB/s both ways; the host side of the
                # box saturates at 110 / 212 / 184 GB/s of pinned uploads for 2 / 4 / 8 ranks, tools/pcie_probe_multi.py)
                "host_link_gbs_per_gpu": (nb * res["n_in"] * 4 + nb * res["n_out"] * 4) / res["e2e_s"] / 1e9,
                "efficiency_vs_one_gpu_824k": (units / res["e2e_s"]) / 824e3 if C == 8 and M == 256 else None},
        "gpu_launches": int(res["launches"]),
        "timed_as": res.get("timed_as", "K direct launches"),
    }
    if res["e2e16_s"] > 0:
        out["e2e_s16_ingest"] = {"value": world * units / res["e2e16_s"], "unit": UNIT, "h2d_bytes_per_step": int(nb * res["n_in"] * 2),
                                 "d2h_bytes_per_step": int(nb * res["n_out"] * 4), "ms_per_step": res["e2e16_s"] * 1e3,
                                 "note": "same call with 16-bit PCM host buffers (btkb200_chain_batch_pcm), converted on the device"}
    if res["adaptive_s"] > 0:
        out["e2e_adaptive_mvdr"] = {
            "value": world * units / res["adaptive_s"], "unit": UNIT, "ms_per_step": res["adaptive_s"] * 1e3,
            "h2d_bytes_per_step": int(nb * res["n_in"] * 4), "d2h_bytes_per_step": int(nb * res["n_out"] * 4),
            "note": "btkb200_mvdr_chain_batch: per utterance covariance (x x^H, ff 0.99) on the first second -> load 1e-2 "
                    "trace/C -> per-bin MVDR solve -> fused chain with that utterance's weights; host buffers in and out"}
    if res["pf_s"] > 0:
        out["e2e_zelinski_postfilter"] = {
            "value": world * units / res["pf_s"], "unit": UNIT, "ms_per_step": res["pf_s"] * 1e3,
            "note": "btkb200_chain_zelinski_batch (analysis -> SubbandDS -> ZelinskiPostFilter alpha 0.6 |.| -> synthesis, "
                    "staged kernels, pipelined over the batch); pinned host buffers in and out"}
    return out


def run_ours(args, wl_cfg, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    old_affinity = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        # NCCL may print its version banner on stdout when the communicator is created; the contract is ONE JSON
        # line on stdout, so fd 1 points at stderr until the first collective has run.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "of fallback (B200_PROFILING.md)"
    traffic_all = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic_all = json.load(open(tpath))
        except Exception:
            traffic_all = {}

    def max_over_ranks(res):
        keys = ["total_ms", "kern_ms", "e2e_s", "e2e16_s", "adaptive_s", "pf_s"]
        tt = torch.tensor([res[k] for k in keys], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        for k, v in zip(keys, tt.tolist()):
            res[k] = float(v)
        return res

    # ---- primary configuration (weak scaling: fixed batch per GPU)
    config_id = 2
    prim = measure(wl_cfg, "ds", wl_cfg["batch"], args, rank, world, local_rank, dev, config_id, args.steps,
                   full_host_inputs=True, with_s16=True, adaptive=args.mvdr, sampler=ClockSampler(local_rank))
    prim = max_over_ranks(prim)

    # ---- the other BASELINE configurations (strong scaling: the utterances are split over the ranks)
    extras = []
    if not args.only_primary and args.workload == "cfg2":
        for key, wl_name, mode, total in EXTRA:
            cfg = dict(WORKLOADS[wl_name])
            nb = max(1, total // world)
            steps = 3 if total >= 512 else max(3, min(args.steps, 10))
            res = measure(cfg, mode, nb, args, rank, world, local_rank, dev, 3 if wl_name == "cfg3" else 4, steps,
                          full_host_inputs=False, with_s16=False, adaptive=(mode == "mvdr"))
            extras.append((key, cfg, mode, steps, max_over_ranks(res)))
    if old_affinity:
        os.sched_setaffinity(0, old_affinity)

    if rank == 0:
        line = {"metric": METRIC, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(wl_cfg, args)}
        line.update(summarise(wl_cfg, "ds", prim, world, args.steps, peak, peak_src, traffic_all.get(args.workload)))
        line["clocks"] = prim["clocks"]
        jobs = [(wl_cfg, "ds", prim["first"], config_id)]
        line["configs"] = {}
        for key, cfg, mode, steps, res in extras:
            sub = summarise(cfg, mode, res, world, steps, peak, peak_src, traffic_all.get(key))
            sub["scaling"] = "strong"
            sub["steps"] = steps
            sub["workload"] = cfg["desc"]
            line["configs"][key] = sub
            jobs.append((cfg, mode, res["first"], 3 if cfg["C"] == 16 else 4))
        if not args.no_cpu_baseline:
            # parity of the timed outputs (first recording of every configuration) against the reference's CPU chain, and
            # the CPU baseline itself (N = 1 only), on the host cores of this box
            cores = os.cpu_count() or 1
            ctx = mp.get_context("spawn")
            with ctx.Pool(min(cores, max(2, len(jobs)))) as pool:
                refs = pool.map(_reference_output_worker, jobs)
            outs0 = [prim["out0"]] + [e[4]["out0"] for e in extras]
            names = ["primary"] + [e[0] for e in extras]
            for name, out0, (ref, kind) in zip(names, outs0, refs):
                par = {"snr_db": snr_db(out0, ref), "against": "oracle/_ref (compiled reference chain)" if kind == "reference"
                       else "oracle port", "gate_db": 70.0, "recording": "first utterance of rank 0, whole length"}
                if par["snr_db"] < 70.0:
                    raise SystemExit(f"bench.py: {name}: timed output is {par['snr_db']:.1f} dB from the reference (< 70 dB)")
                (line if name == "primary" else line["configs"][name])["parity"] = par
            if world == 1:
                with ctx.Pool(cores) as pool:
                    secs = wl_cfg["seconds"]
                    rate, kind, wall = cpu_chain_rate(wl_cfg, secs, 2, cores, pool)
                line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                                        "sample": f"{cores} single-threaded processes x 2 recordings of {wl_cfg['C']} ch x {secs:g} s "
                                                  f"(same chain, same prototype; the FFT under the reference's gsl calls is "
                                                  f"oracle/gsl_shim's radix-2; {wall:.1f} s wall)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="utterances per GPU (default: per workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time K direct launches (an event pair each) instead of one CUDA graph of K launches")
    ap.add_argument("--postfilter", action="store_true", help="also time the chain with the Zelinski post-filter end to end")
    ap.add_argument("--mvdr", action="store_true", help="also time the per-utterance adaptive MVDR path end to end on the primary workload")
    ap.add_argument("--only-primary", action="store_true", help="skip the cfg3 / cfg4 (MVDR, 64-channel) configurations")
    args = ap.parse_args()
    wl_cfg = dict(WORKLOADS[args.workload])
    if args.batch > 0:
        wl_cfg["batch"] = args.batch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, wl_cfg, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, wl_cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
