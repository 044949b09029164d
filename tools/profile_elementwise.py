"""Which autograd nodes launch the small elementwise kernels of the training step? (dev tool)
One eager fwd+bwd of the bench decoder under torch.profiler; prints, per top-level autograd node / forward op,
the elementwise CUDA kernels it launched."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from graph_detr4d_b200 import synthetic as syn, modules
from graph_detr4d_b200.glue import DeferredWgrad

dev = torch.device("cuda")
model = bench.build_model(1, "f32", dev)
feats = [f.to(dev).requires_grad_(True) for f in syn.make_feats(1, 6, 256, syn.LEVEL_SHAPES_928x1600)]
metas = syn.make_img_metas(1, 1)

def step():
    modules.clear_pack_cache()
    for p in model.parameters(): p.grad = None
    st, _, refs = model(feats, metas, 1)
    loss = bench.loss_fn(st, refs)
    with DeferredWgrad() as wq:
        loss.backward()
        wq.flush()

for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
ev = prof.events()
SMALL = {"aten::add", "aten::add_", "aten::mul", "aten::mul_", "aten::threshold_backward", "aten::copy_", "aten::fill_",
         "aten::zero_", "aten::zeros", "aten::sigmoid", "aten::sum", "aten::mean", "aten::div", "aten::neg", "aten::sub",
         "aten::clone", "aten::contiguous", "aten::_foreach_copy_", "aten::cat", "aten::stack", "aten::where",
         "aten::sigmoid_backward", "aten::new_zeros", "aten::zeros_like", "aten::empty_like"}
cnt = collections.Counter()
for e in ev:
    if e.device_type != torch.autograd.DeviceType.CPU or e.name not in SMALL:
        continue
    # skip ops nested inside another SMALL op (count the outermost)
    p, nested, auto, fwd_top = e.cpu_parent, False, None, None
    while p is not None:
        if p.name in SMALL:
            nested = True
        if p.name.startswith("autograd::engine::evaluate_function"):
            auto = p.name.replace("autograd::engine::evaluate_function: ", "")
        fwd_top = p.name
        p = p.cpu_parent
    if nested:
        continue
    if e.name in ("aten::empty_like",):
        continue
    cuda_us = sum(k.duration for k in e.kernels) if e.kernels else 0
    cnt[(auto or "fwd:" + str(fwd_top), e.name, str(e.input_shapes)[:70])] += 1
tot = sum(cnt.values())
print("small torch ops per eager step:", tot)
for (node, op, shp), n in sorted(cnt.items(), key=lambda kv: -kv[1])[:60]:
    print(f"{n:4d}  {node[:42]:42s} {op[:26]:26s} {shp}")
