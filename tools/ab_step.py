"""Resident graphed decoder step time, for A/B of glue changes (dev tool): prints the median of
5 timings of 50 replays.  GD4D_FUSED_GLUE=0 python tools/ab_step.py  -> eager op chains."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from graph_detr4d_b200 import synthetic as syn, fused
from graph_detr4d_b200.graphed import GraphedTrainStep

dev = torch.device("cuda")
T = int(os.environ.get("T", "1"))
model = bench.build_model(T, os.environ.get("DTYPE", "f32"), dev)
feats = [f.to(dev) for f in syn.make_feats(1, 6 * T, 256, syn.LEVEL_SHAPES_928x1600)]
metas = syn.make_img_metas(1, T)
stepper = GraphedTrainStep(model, lambda f: bench.loss_fn(*(lambda st, _, r: (st, r))(*model(f, metas, 1))), feats, metas)
for _ in range(10): stepper.step()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50): stepper.step()
    e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e) / 50)
ts.sort()
print(f"fused_glue={fused.ENABLED} T={T} ms/step median {ts[2]:.3f} min {ts[0]:.3f} max {ts[-1]:.3f}")
