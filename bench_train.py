#!/usr/bin/env python
"""bench_train.py -- BASELINE.json configs[3]: Graph-DETR4D end-to-end training step
(backbone + FPN + 6-layer decoder with the fused cross-view attention) on synthetic
multi-view frames, data-parallel over N GPUs of one node.  Reports train frames/s
(a frame = one scene of 6*T camera images) as ONE JSON line.

Everything outside the decoder is stock library code and deliberately so (SURVEY.md
section 2 marks backbone / neck / loss OUT OF SCOPE): torchvision ResNet-50 (frozen
BN, channels_last, bf16 autocast like the reference's fp16 backbone, DET:68) and a
plain FPN with an extra stride-2 level ('on_output').  The decoder runs fp32 on fp32
features exactly as in the reference (auto_fp16(out_fp32=True)).  Because the FPN
emits channels_last maps the fused kernels read them ZERO-COPY (no pack pass).  The
loss is scaffolding (L1 / BCE against fixed synthetic targets for every query, no
Hungarian matching and therefore no host sync).

  python bench_train.py [--frames T] [--steps K] [--warmup W]
  torchrun --nproc-per-node N bench_train.py --gpus N
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402


class FPN4(nn.Module):
    """C3..C5 -> 256-ch P3..P5 + one extra stride-2 conv on P5 (mmdet FPN, start_level=1,
    add_extra_convs='on_output', num_outs=4: projects/configs/detr3d/detr3d_res50.py:42-49)."""

    def __init__(self, in_channels=(512, 1024, 2048), out_channels=256):
        super().__init__()
        self.lateral = nn.ModuleList([nn.Conv2d(c, out_channels, 1) for c in in_channels])
        self.output = nn.ModuleList([nn.Conv2d(out_channels, out_channels, 3, padding=1) for _ in in_channels])
        self.extra = nn.Conv2d(out_channels, out_channels, 3, stride=2, padding=1)

    def forward(self, feats):
        lat = [l(f) for l, f in zip(self.lateral, feats)]
        for i in range(len(lat) - 1, 0, -1):
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[-2:], mode="nearest")
        outs = [o(x) for o, x in zip(self.output, lat)]
        outs.append(self.extra(outs[-1]))
        return outs


class Detector(nn.Module):
    def __init__(self, num_cams, num_query=900, num_classes=10, code_size=10):
        super().__init__()
        import torchvision
        from graph_detr4d_b200 import synthetic as syn
        from graph_detr4d_b200.decoder import Detr3DTransformer, Detr3DTransformerDecoder
        r = torchvision.models.resnet50(weights=None)
        self.stem = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool)
        self.layers = nn.ModuleList([r.layer1, r.layer2, r.layer3, r.layer4])
        self.neck = FPN4()
        cfg = dict(type="Deform3DCrossAttn", embed_dims=256, num_heads=8, num_levels=4, num_points=4,
                   num_cams=num_cams, pc_range=syn.PC_RANGE, dropout=0.1)
        dec = Detr3DTransformerDecoder(cfg, num_layers=6, dropout=0.1)
        self.transformer = Detr3DTransformer(dec, num_query=num_query, code_size=code_size)
        for i, layer in enumerate(dec.layers):
            syn.randomize_generators(layer.attentions[1], seed=100 + i)
        self.cls_branches = nn.ModuleList([nn.Linear(256, num_classes) for _ in range(6)])
        for m in self.modules():                         # norm_eval=True, frozen BN statistics
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
                for p in m.parameters():
                    p.requires_grad_(False)
        for p in list(self.stem.parameters()) + list(self.layers[0].parameters()):
            p.requires_grad_(False)                      # frozen_stages=1

    def train(self, mode=True):
        super().train(mode)
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
        return self

    def extract_feat(self, img):
        B, N, C, H, W = img.shape
        x = img.view(B * N, C, H, W).contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            x = self.stem(x)
            feats = []
            for i, layer in enumerate(self.layers):
                x = layer(x)
                if i >= 1:
                    feats.append(x)
            outs = self.neck(feats)
        return [o.float().unflatten(0, (B, N)) for o in outs]      # fp32, channels_last strides kept

    def forward(self, img, img_metas):
        feats = self.extract_feat(img)
        states, ref0, refs = self.transformer(feats, img_metas, img.shape[0])
        cls = torch.stack([b(states[i]) for i, b in enumerate(self.cls_branches)])
        reg = torch.stack([b(states[i]) for i, b in enumerate(self.transformer.reg_branches)])  # HD:133-156
        return states, refs, cls, reg


def run_train(world, rank, local_rank, dev, T=2, steps=10, warmup=3):
    """Time the end-to-end training step; returns the result dict on rank 0 (None elsewhere).
    The caller owns process-group setup (bench.py calls this as one of its extra legs)."""
    torch.backends.cudnn.benchmark = True
    from graph_detr4d_b200 import _lib, ops, synthetic as syn
    _lib.load(build_if_missing=False)
    N = 6 * T
    torch.manual_seed(0)
    model = Detector(N).to(dev).train()
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], broadcast_buffers=False,
                                                        gradient_as_bucket_view=True)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=2e-4, weight_decay=0.01,
                            fused=True)
    g = torch.Generator().manual_seed(rank)
    img_host = torch.randn(1, N, 3, 928, 1600, generator=g).pin_memory()
    metas = syn.make_img_metas(1, T)
    tgt_cls = (torch.rand(6, 900, 1, 10, generator=g) > 0.9).float().to(dev)
    tgt_w = torch.randn(6, 900, 1, 256, generator=g).to(dev)

    def step():
        img = img_host.to(dev, non_blocking=True)                 # H2D of the step's 6T images
        states, refs, cls, reg = net(img, metas)
        loss = F.binary_cross_entropy_with_logits(cls, tgt_cls) + 0.25 * reg.abs().mean() + (states * tgt_w).mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad], 35.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        step()
    calls0 = ops.launch_count()
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        loss = step()
    e.record()
    barrier()
    ms = s.elapsed_time(e)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    line = None
    if rank == 0:
        line = dict(metric="train_frames_per_s", value=world * steps / (ms * 1e-3), unit="frames/s",
                    n_gpus=world, steps=steps, warmup=max(warmup, 3), ms_per_step=ms / steps,
                    higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="bf16 backbone+FPN (autocast) / f32 decoder", data="synthetic",
                    config=dict(workload=f"graph_detr4d_e2e_train_T{T}_N{N}_928x1600_res50_fpn_dec6_Q900",
                                per_gpu_batch=1, parallelism=f"ddp{world}", loss="synthetic (no Hungarian)",
                                features="FPN emits channels_last -> fused kernels read zero-copy",
                                h2d_bytes_per_step=img_host.numel() * 4),
                    gpu_launches=(ops.launch_count() - calls0), final_loss=float(loss),
                    peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2))
    del net, model, opt, img_host
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=2)
    args = ap.parse_args()
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    line = run_train(world, rank, local_rank, dev, T=args.frames, steps=args.steps, warmup=args.warmup)
    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
