"""Exhaustive shape sweep of the device path against the numpy oracle: every (M, m, r) the library accepts x two delay
compensation types, random channel counts and weights; fused chain, staged analysis / beamform / synthesis.
usage (GPU box): python tools/gpu_full_sweep.py"""
import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import btk_b200, btk_oracle as bo
wl = btk_b200.workloads
bad = 0; n = 0; unsup = []
for M in (64, 128, 256, 512, 1024):
    for r in (0, 1, 2, 3):
        for m in (1, 2, 3, 4):
            for dct in (0, 2):
                if dct == 2 and m * (1 << r) < 2:
                    continue        # refused by plan_create (undefined in the reference)
                rng = np.random.default_rng(M * 1000 + m * 100 + r * 10 + dct)
                C = int(rng.choice([1, 2, 3, 5, 8])); D = M >> r
                h, g = wl.kaiser_prototype(M, m, r)
                geo = bo.BankGeometry(M, m, r, dct)
                try:
                    plan = btk_b200.Plan(M, m, r, C, h, g, dct=dct)
                except btk_b200.BtkError as e:
                    unsup.append((M, m, r)); continue
                W = (rng.standard_normal((geo.B, C)) + 1j * rng.standard_normal((geo.B, C))) / C
                plan.set_weights(W)
                T = int(rng.integers(20 * D, 30 * D)) + int(rng.integers(0, D))
                pcm = wl.noise_recording(T, C, seed=int(rng.integers(1 << 30)), sigma=700.0)
                X, Y, ref = bo.chain(pcm, h, g, geo, W)
                out = plan.chain(pcm)
                snap = plan.analysis(pcm)
                e1 = bo.rel_l2(snap, X[:, :, :geo.B].transpose(0, 2, 1))
                Yd = plan.beamform(snap)
                e2 = bo.rel_l2(Yd, Y[:, :geo.B])
                s1 = bo.snr_db(out, ref) if out.shape == ref.shape else -1
                syn = plan.synthesis(Yd)
                s2 = bo.snr_db(syn, ref) if syn.shape == ref.shape else -2
                if s1 < 0 or s2 < 0: print('SHAPES', (M, m, r, dct, C, T), out.shape, syn.shape, ref.shape, snap.shape, X.shape, flush=True)
                n += 1
                if not (s1 >= 70 and s2 >= 70 and e1 <= 1e-4 and e2 <= 1e-4):
                    bad += 1; print("BAD", (M, m, r, dct, C, T), s1, s2, e1, e2, flush=True)
                plan.close()
print("checked", n, "bad", bad, "unsupported", sorted(set(unsup)))
