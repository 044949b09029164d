"""In-process A/B of the fused sampling kernels with / without GD4D_FLAG_L2_PREFETCH at the flagship
shapes: prepared launches of both variants timed interleaved (CUDA events, rotating 3 value-map
copies so consecutive launches do not reuse L2).  (dev tool)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graph_detr4d_b200 import ops, synthetic as syn
from graph_detr4d_b200.ops import MODE_C, XViewConfig
from tests import helpers as H

res = {}
for name, T, dtype, wide in [("C_T1_fp32_wide", 1, torch.float32, True), ("C_T2_fp32_wide", 2, torch.float32, True),
                             ("C_T2_bf16_wide", 2, torch.bfloat16, True), ("C_T2_fp32_narrow", 2, torch.float32, False)]:
    sc = H.scene(B=1, T=T, Q=900, shapes=H.FULL_SHAPES)
    packed = ops.pack_features([f.cuda() for f in sc["feats"]], dtype)
    sets = [[v.clone() for v in packed.levels] for _ in range(3)]
    ref, l2i = sc["ref"].cuda(), sc["l2i"].cuda()
    logits, offsets, cam = (t.cuda() for t in H.rand_inputs_c(sc, off_std=1.5))
    cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=wide)
    out, _ = ops.xview_forward(cfg, sets[0], 1, sc["N"], ref, logits, offsets, cam, l2i)
    o0 = out[0] if wide else out
    gout = torch.randn_like(o0)
    gws = torch.randn_like(out[1]) if wide else None
    gsets = [[torch.zeros(v.shape, device="cuda", dtype=torch.float32) for v in s] for s in sets]
    prep = {}
    for flag in (False, True):
        ops.L2_PREFETCH = flag
        prep[flag] = ([ops.prepare_forward(cfg, s, 1, sc["N"], ref, logits, offsets, cam, l2i) for s in sets],
                      [ops.prepare_backward(cfg, s, 1, sc["N"], ref, logits, offsets, cam, l2i, gout, g, grad_wsum=gws)
                       for s, g in zip(sets, gsets)])
    ops.L2_PREFETCH = False
    same = torch.equal(prep[False][0][0].out, prep[True][0][0].out) if False else None

    def t(launches, n=60):
        for i in range(6): launches[i % 3].launch()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(n): launches[i % 3].launch()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / n * 1e3
    r = dict(fwd_off=[], fwd_on=[], bwd_off=[], bwd_on=[])
    for _ in range(5):
        r["fwd_off"].append(t(prep[False][0])); r["fwd_on"].append(t(prep[True][0]))
        r["bwd_off"].append(t(prep[False][1])); r["bwd_on"].append(t(prep[True][1]))
    prep[False][0][0].launch(); prep[True][0][1].launch(); torch.cuda.synchronize()
    res[name] = {k: round(sorted(v)[2], 1) for k, v in r.items()}
    print(name, res[name], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/ab_l2_prefetch.json", "w"), indent=1)
