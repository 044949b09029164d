"""Launch each fused kernel a few times at the flagship shapes (for `ncu --set full`).
Order of launches (REPS each): per case  fwd, bwd (atomics), and for the wide cases bwdS (sorted: 5 kernels).
Cases: C-wide fp32 N=6 | C-wide fp32 N=12 | C-wide bf16 N=12 | C-narrow fp32 N=12 | A fp32 N=6.
tools/summarize_profiles.py full ... <the printed CASES line> names the launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graph_detr4d_b200 import ops, synthetic as syn
from graph_detr4d_b200.ops import MODE_A, MODE_C, XViewConfig
from tests import helpers as H

REPS = int(os.environ.get("REPS", "2"))
cases = [("Cw_f32_N6", MODE_C, 1, torch.float32, True), ("Cw_f32_N12", MODE_C, 2, torch.float32, True),
         ("Cw_bf16_N12", MODE_C, 2, torch.bfloat16, True), ("Cn_f32_N12", MODE_C, 2, torch.float32, False),
         ("A_f32_N6", MODE_A, 1, torch.float32, False)]
only = os.environ.get("CASES")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for name, mode, T, dtype, wide in cases:
    if only and name not in only.split(","):
        continue
    sc = H.scene(B=1, T=T, Q=900, shapes=H.FULL_SHAPES)
    packed = ops.pack_features([f.cuda() for f in sc["feats"]], dtype)
    ref, l2i = sc["ref"].cuda(), sc["l2i"].cuda()
    if mode == MODE_C:
        logits, offsets, cam = (t.cuda() for t in H.rand_inputs_c(sc))
        cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=wide)
    else:
        logits, offsets, cam = H.rand_inputs_a(sc).cuda(), None, None
        cfg = XViewConfig(MODE_A, 8, 1, tuple(syn.PC_RANGE), 900.0, 1600.0)
    f = ops.prepare_forward(cfg, packed.levels, 1, sc["N"], ref, logits, offsets, cam, l2i)
    gout = torch.randn_like(f.out)
    gws = torch.randn_like(f.wsum) if f.wsum is not None else None
    gv = [torch.zeros(v.shape, device="cuda", dtype=torch.float32) for v in packed.levels]
    ops.SORTED_BACKWARD = False
    b = ops.prepare_backward(cfg, packed.levels, 1, sc["N"], ref, logits, offsets, cam, l2i, gout, gv, gws)
    bs = None
    if wide:
        ops.SORTED_BACKWARD = True
        bs = ops.prepare_backward(cfg, packed.levels, 1, sc["N"], ref, logits, offsets, cam, l2i, gout, gv, gws)
    ops.SORTED_BACKWARD = "auto"
    for _ in range(REPS):
        flush.zero_()
        f.launch()
        flush.zero_()
        b.launch()
        if bs is not None:
            flush.zero_()
            bs.launch()
    torch.cuda.synchronize()
    print(name, "done", flush=True)
print("CASES", " ".join(sum(([n + "_fwd", n + "_bwd"] + ([n + "_bwdS"] if w else []) for n, _, _, _, w in cases
                             if not only or n in only.split(",")), [])))
