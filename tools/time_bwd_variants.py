"""A/B of the two wide backward implementations on the bench's own layer-0 inputs (dev tool):
atomics-per-corner (xview_bwd.cu) vs sorted owner-computes (xview_bwd_sorted.cu).
  python tools/time_bwd_variants.py [--reps 200] [--once]     (--once: a single call of each, for ncu)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from graph_detr4d_b200 import ops, synthetic as syn

reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 200
once = "--once" in sys.argv
dev = torch.device("cuda", 0)
out = {}
for T, dtype in ((1, "f32"), (2, "f32"), (2, "bf16")):
    model = bench.build_model(T, dtype, dev, seed=0)
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    feats = [f.to(dev, tdt) for f in syn.make_feats(1, 6 * T, 256, syn.LEVEL_SHAPES_928x1600, seed=0)]
    metas = syn.make_img_metas(1, T)
    for srt in (False, True):
        ops.SORTED_BACKWARD = srt
        rb, rf = bench.kernel_roofline(model, feats, metas, T, dtype, dev, reps=1 if once else reps)
        out[f"T{T}_{dtype}_{'sorted' if srt else 'atomics'}"] = dict(bwd_us=rb["us_per_launch"], fwd_us=rf["us_per_launch"])
        print(T, dtype, "sorted" if srt else "atomics", rb["us_per_launch"], flush=True)
    del model, feats
    torch.cuda.empty_cache()
    if once:
        break
ops.SORTED_BACKWARD = "auto"
print(json.dumps(out))
