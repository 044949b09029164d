"""Resident graphed decoder step time for A/B of glue changes (dev tool).  All variants are
captured in ONE process and timed interleaved (round-robin, 8 rounds x 40 replays), because
process-to-process and minute-to-minute drift on a shared B200 is ~4 % -- larger than most of
the effects being measured.   VARIANTS=base,no_adamw,no_gen,no_fused python tools/ab_step.py"""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from graph_detr4d_b200 import synthetic as syn, fused, graphed, modules, ops
from graph_detr4d_b200.graphed import GraphedTrainStep

dev = torch.device("cuda")
T = int(os.environ.get("T", "1"))
names = os.environ.get("VARIANTS", "base,no_adamw,no_gen,no_fused").split(",")
feats = [f.to(dev) for f in syn.make_feats(1, 6 * T, 256, syn.LEVEL_SHAPES_928x1600)]
metas = syn.make_img_metas(1, T)
base_model = bench.build_model(T, os.environ.get("DTYPE", "f32"), dev)


def build(name):
    fused.ENABLED = name != "no_fused"
    modules._PACKED_GEN = name != "no_gen"
    graphed.MULTI_TENSOR_ADAMW = name != "no_adamw"
    fused.SOFTMAX_BWD = name != "no_smbwd"
    ops.SORTED_BACKWARD = {"sorted": True, "atomics": False}.get(name, "auto")
    ops.PRESORT = name == "presort"
    ops.FWD_EMIT = name != "no_fwd_emit"
    fused.LN_COPIES = name != "no_copies"
    model = copy.deepcopy(base_model)
    st = GraphedTrainStep(model, lambda f: bench.loss_fn(*(lambda s, _, r: (s, r))(*model(f, metas, 1))), feats, metas)
    fused.ENABLED, modules._PACKED_GEN, graphed.MULTI_TENSOR_ADAMW, fused.SOFTMAX_BWD = True, True, True, True
    ops.SORTED_BACKWARD = "auto"
    ops.PRESORT = False
    ops.FWD_EMIT = True
    fused.LN_COPIES = True
    return st


steppers = {n: build(n) for n in names}
for st in steppers.values():
    for _ in range(10): st.step()
torch.cuda.synchronize()
times = {n: [] for n in names}
for _ in range(8):
    for n, st in steppers.items():
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(40): st.step()
        e.record(); torch.cuda.synchronize()
        times[n].append(s.elapsed_time(e) / 40)
for n in names:
    ts = sorted(times[n])
    print(f"{n:10s} T={T} ms/step median {ts[len(ts)//2]:.3f} min {ts[0]:.3f} max {ts[-1]:.3f}")
