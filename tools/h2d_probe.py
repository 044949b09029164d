"""How much pinned-host -> device bandwidth does each rank get when N ranks copy at once?  (dev tool)
torchrun --nproc-per-node N tools/h2d_probe.py  -> one JSON line from rank 0: per-rank GB/s with every
rank copying concurrently, and rank 0 copying alone.  189 MB buffer = one step's fp32 feature maps."""
import json, os, sys, time
import torch, torch.distributed as dist

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nbytes = int(os.environ.get("BYTES", 189388800))
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
host.fill_(1)
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
back = torch.empty(nbytes, dtype=torch.uint8).pin_memory()

def rate(n=20, src=host, dst=dev):
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[lr]); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        dst.copy_(src, non_blocking=True)
    e.record(); torch.cuda.synchronize()
    return n * nbytes / (s.elapsed_time(e) * 1e-3) / 1e9

allr = rate()
d2h = rate(src=dev, dst=back)
if world > 1:
    t = torch.tensor([allr, d2h], device="cuda"); outs = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    per = [[round(float(x), 1) for x in o] for o in outs]
    dist.barrier(device_ids=[lr]); torch.cuda.synchronize()
    alone = None
    if rank == 0:
        for _ in range(2): dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10): dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        alone = 10 * nbytes / (time.perf_counter() - t0) / 1e9
    dist.barrier(device_ids=[lr])
else:
    per, alone = [[round(allr, 1), round(d2h, 1)]], allr
if rank == 0:
    numa = sorted(os.listdir("/sys/devices/system/node")) if os.path.isdir("/sys/devices/system/node") else None
    print(json.dumps(dict(world=world, bytes=nbytes, h2d_d2h_GBps_per_rank_concurrent=per,
                          h2d_total_GBps=round(sum(p[0] for p in per), 1), rank0_alone_GBps=alone and round(alone, 1),
                          cpus=len(os.sched_getaffinity(0)), numa_nodes=[n for n in (numa or []) if n.startswith("node")])))
if world > 1:
    dist.destroy_process_group()
