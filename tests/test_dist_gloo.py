"""World-size-2 CPU (gloo) test of the data-parallel host logic: the decoder built
around the oracle port trains under the same flat-gradient / single all-reduce scheme
GraphedTrainStep uses on GPUs, scenes are sharded one per rank with no data-path
collective, and the step time is reduced as a MAX over ranks."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import warnings
    warnings.filterwarnings("ignore")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from graph_detr4d_b200 import synthetic as syn
    from graph_detr4d_b200.decoder import Detr3DTransformer, Detr3DTransformerDecoder
    from oracle.modules_port import build_oracle_attention
    from tests import helpers as H

    torch.manual_seed(0)                                   # identical replicas
    cfg = dict(type="Deform3DCrossAttn", embed_dims=64, num_heads=2, num_levels=4, num_points=2,
               num_cams=6, pc_range=syn.PC_RANGE, dropout=0.0)
    dec = Detr3DTransformerDecoder(cfg, num_layers=2, embed_dims=64, num_heads=2, feedforward_channels=128,
                                   dropout=0.0, cross_attn_factory=build_oracle_attention)
    model = Detr3DTransformer(dec, num_query=32)
    for i, layer in enumerate(dec.layers):
        syn.randomize_generators(layer.attentions[1], seed=40 + i)
    params = [p for p in model.parameters()]
    flat = torch.zeros(sum(p.numel() for p in params))
    o = 0
    for p in params:
        p.grad = flat[o:o + p.numel()].view_as(p)
        o += p.numel()
    # one scene per rank (different seed): weak scaling, no exchange on the data path
    feats = syn.make_feats(1, 6, 64, [(12, 20), (6, 10), (3, 5), (2, 3)], seed=rank)
    metas = syn.make_img_metas(1, 1)
    states, _, refs = model(feats, metas, 1)
    gout = torch.randn(states.shape, generator=torch.Generator().manual_seed(3))
    (states * gout).sum().backward()
    local = flat.clone()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    # grouped in-place all-reduce of several buffers (GraphedTrainStep's N > 1 scheme) == one by one
    bufs = [torch.full((3, 4), float(rank + 1)), torch.arange(5.0) * (rank + 1), torch.ones(2, 2, 2) * rank]
    want = [b.clone() for b in bufs]
    for w in want:
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    with dist.distributed_c10d._coalescing_manager():
        for b in bufs:
            dist.all_reduce(b, op=dist.ReduceOp.SUM)
    coalesced_ok = all(torch.equal(a, b) for a, b in zip(bufs, want))
    t = torch.tensor([10.0 + rank])                        # max-over-ranks timing reduction
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        torch.save(dict(avg=flat.clone(), locals=gathered, tmax=float(t), coalesced_ok=coalesced_ok), out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_allreduce(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["tmax"] == 11.0 and r["coalesced_ok"]
    l0, l1 = r["locals"]
    assert not torch.equal(l0, l1)                         # ranks saw different scenes
    assert torch.allclose(r["avg"], (l0 + l1) / 2, rtol=0, atol=1e-6)
    assert r["avg"].abs().max() > 0
