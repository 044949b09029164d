"""Scratch timing of the fused kernels at the flagship shapes (dev tool, not the bench)."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graph_detr4d_b200 import ops, synthetic as syn
from graph_detr4d_b200.ops import MODE_A, MODE_C, XViewConfig
from tests import helpers as H

def timeit(fn, iters=50, warm=5, flush=None):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None: flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts)//2], ts[0]

res = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
import sys as _s
WIDE = "--narrow" not in _s.argv
for name, mode, T, dtype in [("C_T1_fp32", MODE_C, 1, torch.float32), ("C_T2_fp32", MODE_C, 2, torch.float32),
                             ("C_T2_bf16", MODE_C, 2, torch.bfloat16), ("A_T1_fp32", MODE_A, 1, torch.float32)]:
    sc = H.scene(B=1, T=T, Q=900, shapes=H.FULL_SHAPES)
    t0 = time.time()
    packed = ops.pack_features([f.cuda() for f in sc["feats"]], dtype)
    ref = sc["ref"].cuda(); l2i = sc["l2i"].cuda()
    if mode == MODE_C:
        logits, offsets, cam = (t.cuda() for t in H.rand_inputs_c(sc))
        cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=WIDE)
        fwd = lambda: ops.xview_forward(cfg, packed.levels, 1, sc["N"], ref, logits, offsets, cam, l2i)
    else:
        logits = H.rand_inputs_a(sc).cuda(); offsets = cam = None
        cfg = XViewConfig(MODE_A, 8, 1, tuple(syn.PC_RANGE), 900.0, 1600.0)
        fwd = lambda: ops.xview_forward(cfg, packed.levels, 1, sc["N"], ref, logits, lidar2img=l2i)
    out, mask = ops.xview_forward(cfg, packed.levels, 1, sc["N"], ref, logits, offsets, cam, l2i, want_mask=True)
    if isinstance(out, tuple): out = out[0]
    gout = torch.randn_like(out)
    gvals = [torch.zeros(v.shape, device='cuda', dtype=torch.float32) for v in packed.levels]
    if "--static" in _s.argv: ops.DYNAMIC_SCHEDULE = False
    if "--tma" in _s.argv: ops.TMA_FORWARD = True
    fprep = ops.prepare_forward(cfg, packed.levels, 1, sc["N"], ref, logits, offsets, cam, l2i)
    fwd = fprep.launch
    bprep = ops.prepare_backward(cfg, packed.levels, 1, sc["N"], ref, logits, offsets, cam, l2i, gout, gvals)
    bwd = bprep.launch
    if "--no-gv" in _s.argv:
        for l in range(len(gvals)): bprep.params.grad_value[l] = None
    feats_gpu = [f.cuda() for f in sc["feats"]]
    pk = lambda: ops.pack_features(feats_gpu, dtype)
    res[name] = dict(valid_frac=float(mask.float().mean()),
                     fwd_ms_cold=timeit(fwd, flush=flush), fwd_ms_warm=timeit(fwd),
                     bwd_ms_cold=timeit(bwd, flush=flush), bwd_ms_warm=timeit(bwd),
                     pack_ms=timeit(pk, iters=10))
    print(name, res[name], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/quick_time.json", "w"), indent=1)
