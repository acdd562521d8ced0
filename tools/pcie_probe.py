"""PCIe ceiling of the end-to-end path: the bench step's upload (491.5 MB pinned) alone and with its download running against it.
usage (GPU box): python tools/pcie_probe.py"""
import torch, time
x = torch.empty(491520000 // 4, dtype=torch.float32, pin_memory=True)
d = torch.empty_like(x, device='cuda')
o = torch.empty(61440000 // 4, dtype=torch.float32, pin_memory=True)
do = torch.empty_like(o, device='cuda')
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t = (time.perf_counter() - t0) / 10
print("H2D alone: %.2f ms  %.1f GB/s" % (t * 1e3, x.numel() * 4 / t / 1e9))
s2 = torch.cuda.Stream()
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): o.copy_(do, non_blocking=True)
torch.cuda.synchronize()
t = (time.perf_counter() - t0) / 10
print("H2D + concurrent D2H: %.2f ms  %.1f GB/s H2D" % (t * 1e3, x.numel() * 4 / t / 1e9))
