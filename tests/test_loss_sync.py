"""Row f3, second half: all layers' loss normalisers (detr3d_head.py:316-331) with one packed all-reduce
and no host sync, vs the reference's per-layer lines restated in oracle/loss_sync_oracle.py --
single process and world size 2 over gloo (each rank holds a different number of ground-truth boxes)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_detr4d_b200 import loss_sync  # noqa: E402
from oracle import loss_sync_oracle as lo  # noqa: E402


def test_hungarian_counts():
    pos, neg = loss_sync.hungarian_pos_neg_counts(6, 900, [37, 0, 1200])
    assert pos.tolist() == [37 + 0 + 900] * 6 and neg.tolist() == [3 * 900 - 937] * 6


@pytest.mark.parametrize("bg,sync", [(0.0, True), (0.1, True), (0.1, False)])
def test_single_process_equals_reference_lines(bg, sync):
    pos, neg = [3, 0, 41, 7, 900, 1], [897, 900, 859, 893, 0, 899]
    cls, npos = loss_sync.packed_avg_factors(pos, neg, bg, sync)
    like = torch.zeros(1)
    for l in range(6):
        c_ref, p_ref = lo.layer_avg_factors(pos[l], neg[l], bg, sync, like)
        # (without sync the reference keeps a python double; the loss divides an fp32 tensor by it, i.e. by its fp32 value)
        assert float(cls[l]) == float(np.float32(float(c_ref))) and float(npos[l]) == float(np.float32(p_ref))
    assert cls.dtype == torch.float32 and tuple(cls.shape) == (6,)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, Q = 6, 900
    gts = [[5, 0], [38, 2]][rank]                              # 2 samples per rank, different gt counts
    pos, neg = loss_sync.hungarian_pos_neg_counts(L, Q, gts)
    pos = pos + np.arange(L) * rank                            # make the layers differ
    neg = neg - np.arange(L) * rank
    calls = []
    real = dist.all_reduce

    def counting(t, *a, **k):
        calls.append(t.numel())
        return real(t, *a, **k)

    dist.all_reduce = counting
    try:
        res = {}
        for bg, sync in ((0.1, True), (0.1, False)):
            calls.clear()
            cls, npos = loss_sync.packed_avg_factors(pos, neg, bg, sync)
            n_packed = list(calls)
            calls.clear()
            ref = [lo.layer_avg_factors(int(pos[l]), int(neg[l]), bg, sync, torch.zeros(1)) for l in range(L)]
            n_ref = list(calls)
            res[(bg, sync)] = dict(cls=cls.tolist(), npos=npos.tolist(),
                                   ref_cls=[float(np.float32(float(c))) for c, _ in ref],
                                   ref_npos=[float(np.float32(p)) for _, p in ref], n_packed=n_packed, n_ref=n_ref)
    finally:
        dist.all_reduce = real
    out[rank] = res
    dist.destroy_process_group()


def test_world2_gloo_one_collective_same_numbers():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for rank in range(world):
        for key, r in out[rank].items():
            assert r["cls"] == r["ref_cls"], (rank, key)          # bit-identical to the per-layer lines
            assert r["npos"] == r["ref_npos"], (rank, key)
            assert len(r["n_packed"]) == 1                        # ONE all-reduce for all 6 layers ...
            assert len(r["n_ref"]) == (12 if key[1] else 6)       # ... where the reference issues 6 or 12
    assert out[0][(0.1, True)]["cls"] == out[1][(0.1, True)]["cls"]      # replicas agree
    # the averaged positive count really mixes the two ranks' ground truth
    assert out[0][(0.1, True)]["npos"][0] == pytest.approx((5 + 0 + 38 + 2) / 2)
