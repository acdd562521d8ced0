"""Aggregate host->device ceiling with one process per GPU copying at the same time (what bounds the 8-rank end-to-end number).
usage (GPU box): python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/pcie_probe_multi.py"""
import os, time, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", lr))
x = torch.empty(491520000 // 4, dtype=torch.float32, pin_memory=True)
d = torch.empty_like(x, device="cuda")
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t = torch.tensor([(time.perf_counter() - t0) / 10], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    tt = float(t.item())
    print("%d ranks: %.2f ms per 491.5 MB upload each, %.1f GB/s per GPU, %.1f GB/s aggregate" % (world, tt * 1e3, 0.49152 / tt, world * 0.49152 / tt))
dist.destroy_process_group()
