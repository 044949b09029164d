"""BASELINE.json configs[4]: cross-view sampling sweep (queries 300-3600, graph neighbours
(points per head) 1-16, frames 1-4, fp32 vs bf16) -- fused kernels alone, HBM GB/s vs roofline.
One factor at a time around (Q=900, P=4, T=2) plus the largest corner.  Writes JSON lines."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graph_detr4d_b200 import ops, roofline, synthetic as syn
from graph_detr4d_b200.ops import MODE_C, XViewConfig

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
dev = "cuda"
feat_cache = {}


def feats(T, dtype):
    key = (T, dtype)
    if key not in feat_cache:
        feat_cache.clear()
        g = torch.Generator(device=dev).manual_seed(T)
        lv = [torch.randn(6 * T, h, w, 256, device=dev, generator=g).to(dtype) for (h, w) in syn.LEVEL_SHAPES_928x1600]
        feat_cache[key] = [[v.clone() for v in lv] for _ in range(2)]      # 2 rotating copies
    return feat_cache[key]


def time_loop(fns, reps=30):
    for i in range(4):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(reps):
        fns[i % len(fns)]()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e-3


def point(Q, P, T, dtype, wide=True):
    N, Hh, L, C = 6 * T, 8, 4, 256
    sets = feats(T, dtype)
    g = torch.Generator().manual_seed(Q * 131 + P * 7 + T)
    ref = torch.rand(1, Q, 3, generator=g).to(dev)
    logits = torch.randn(1, Q, Hh * L * P, generator=g).to(dev)
    offsets = (torch.randn(1, Q, Hh * P * 3, generator=g) * 2.0).to(dev)
    cam = torch.randn(1, Q, N, generator=g).to(dev)
    import numpy as np
    l2i = torch.as_tensor(syn.make_lidar2img(T).astype(np.float32)).unsqueeze(0).to(dev)
    cfg = XViewConfig(MODE_C, Hh, P, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=wide)
    st = roofline.count_corner_reads(MODE_C, syn.LEVEL_SHAPES_928x1600, ref, offsets, l2i, syn.PC_RANGE,
                                     900.0, 1600.0, Hh, P)
    eb = 2 if dtype == torch.bfloat16 else 4
    ab = roofline.algorithmic_bytes(MODE_C, st, 1, Q, N, C, Hh, L, P, eb, wide=wide)
    fw = [ops.prepare_forward(cfg, s, 1, N, ref, logits, offsets, cam, l2i) for s in sets]
    gout = torch.randn_like(fw[0].out)
    gws = torch.randn_like(fw[0].wsum) if wide else None
    gv = [torch.zeros(v.shape, device=dev, dtype=torch.float32) for v in sets[0]]
    bw = [ops.prepare_backward(cfg, s, 1, N, ref, logits, offsets, cam, l2i, gout, gv, gws) for s in sets]
    tf = time_loop([f.launch for f in fw])
    tb = time_loop([b.launch for b in bw])
    return dict(Q=Q, P=P, T=T, N=N, dtype="bf16" if eb == 2 else "f32", mode="wide" if wide else "narrow",
                valid_fraction=round(st["valid_fraction"], 4), corner_reads=ab["S"],
                fwd_us=round(tf * 1e6, 1), bwd_us=round(tb * 1e6, 1),
                fwd_alg_MB=round(ab["fwd"] / 1e6, 1), bwd_alg_MB=round(ab["bwd"] / 1e6, 1),
                fwd_GBs=round(ab["fwd"] / tf / 1e9), bwd_GBs=round(ab["bwd"] / tb / 1e9),
                fwd_frac_of_measured_hbm=round(ab["fwd"] / tf / 1e9 / PEAK, 3),
                bwd_frac_of_measured_hbm=round(ab["bwd"] / tb / 1e9 / PEAK, 3),
                fwd_Mqueries_s=round(Q / tf / 1e6, 2), fwdbwd_Mqueries_s=round(Q / (tf + tb) / 1e6, 2))


if __name__ == "__main__":
    pts = []
    for dtype in (torch.float32, torch.bfloat16):
        for T in (1, 2, 3, 4):
            pts.append((900, 4, T, dtype))
        for Q in (300, 1800, 2700, 3600):
            pts.append((Q, 4, 2, dtype))
        for P in (1, 2, 8, 16):
            pts.append((900, P, 2, dtype))
        pts.append((3600, 16, 4, dtype))
    pts.sort(key=lambda t: (str(t[3]), t[2]))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sweep_r1.jsonl", "w") as f:
        for (Q, P, T, dtype) in pts:
            for wide in (True, False):
                r = point(Q, P, T, dtype, wide)
                f.write(json.dumps(r) + "\n")
                f.flush()
                print(r, flush=True)
