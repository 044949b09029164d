"""torch.profiler (CUPTI) kernel-time breakdown of one eager decoder step (dev tool)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from graph_detr4d_b200 import synthetic as syn, modules
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda")
T = int(os.environ.get("T", "1"))
model = bench.build_model(T, os.environ.get("DTYPE", "f32"), dev)
from graph_detr4d_b200.optim import MultiTensorAdamW
opt = MultiTensorAdamW(list(model.parameters()), lr=2e-4)
feats = [f.to(dev).requires_grad_(True) for f in syn.make_feats(1, 6 * T, 256, syn.LEVEL_SHAPES_928x1600)]
metas = syn.make_img_metas(1, T)

from graph_detr4d_b200.glue import DeferredWgrad
def step():
    modules.clear_pack_cache()
    st, _, refs = model(feats, metas, 1)
    with DeferredWgrad() as wq:
        bench.loss_fn(st, refs).backward(); wq.flush()
    opt.step(); opt.zero_grad(set_to_none=True)
    for f in feats: f.grad = None

for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
from torch.autograd import DeviceType
kern = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA and e.self_device_time_total > 0]
rows = sorted(((e.self_device_time_total / 3.0, e.count / 3, e.key) for e in kern), reverse=True)
tot = sum(r[0] for r in rows)
print(f"kernel time per step: {tot/1e3:.3f} ms in {sum(r[1] for r in rows):.0f} launches")
ours = sum(r[0] for r in rows if "gd4d::" in r[2])
print(f"libgd4d_xview.so kernels: {ours/1e3:.3f} ms ({100*ours/tot:.1f}%), {sum(r[1] for r in rows if 'gd4d::' in r[2]):.0f} launches")
for t, n, k in rows[:70]:
    print(f"{t:9.1f} us {100*t/tot:5.1f}% n={n:6.1f} {k[:150]}")

# ---- which tensors do the elementwise/copy launches touch? (aten op + input shapes) -------------------
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof2:
    for _ in range(3): step()
    torch.cuda.synchronize()
rows = []
for e in prof2.key_averages(group_by_input_shape=True):
    if e.key in ("aten::copy_", "aten::add", "aten::add_", "aten::fill_", "aten::zero_", "aten::cat", "aten::mul",
                 "aten::sum", "aten::stack", "aten::contiguous", "aten::clone") and e.self_device_time_total > 0:
        rows.append((e.self_device_time_total / 3.0, e.count / 3, e.key, str(e.input_shapes)[:120]))
print("\n== elementwise / copy ops by input shape (self device time per step) ==")
for t, n, k, sh in sorted(rows, reverse=True)[:40]:
    print(f"{t:9.1f} us n={n:5.1f} {k:14s} {sh}")
