"""Where does the N>1 step time go?  (dev tool; torchrun --nproc-per-node N tools/dist_breakdown.py)
Times, with CUDA events on rank 0 after a barrier: the fwd/bwd graph alone, the gradient all-reduce alone,
the optimizer graph alone and the full step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from graph_detr4d_b200 import synthetic as syn
from graph_detr4d_b200.graphed import GraphedTrainStep

world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
model = bench.build_model(1, "f32", dev)
feats = [f.to(dev) for f in syn.make_feats(1, 6, 256, syn.LEVEL_SHAPES_928x1600, seed=rank)]
metas = syn.make_img_metas(1, 1)
st = GraphedTrainStep(model, lambda f: bench.loss_fn(*(lambda s, _, r: (s, r))(*model(f, metas, 1))), feats, metas,
                      world_size=world)


def timeit(fn, n=40):
    for _ in range(5): fn()
    if world > 1: dist.barrier(device_ids=[lr])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for _ in range(10): st.step()
torch.cuda.synchronize()
chk = torch.stack([p.detach().double().abs().sum() for p in model.parameters()]).sum().reshape(1)
if world > 1:                                       # replicas must stay identical: same checksum on every rank
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(float(c) == float(allc[0]) for c in allc)
else:
    same = True
if rank == 0:
    print(f"world {world} after 10 steps: param |sum| = {float(chk):.9e}  identical across ranks: {same}", flush=True)
for _ in range(10): st.step()
res = dict(full=timeit(st.step), fb=timeit(st.graph_fb.replay), allreduce=timeit(st._allreduce),
           opt=timeit(st.graph_opt.replay), full2=timeit(st.step))
if rank == 0:
    print("world", world, {k: round(v, 3) for k, v in res.items()}, "flat_grad MB", st.flat_grad.numel() * 4 / 1e6, flush=True)
if world > 1:
    dist.destroy_process_group()
