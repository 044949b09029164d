"""Every product entry point of rows f2-f4 refuses CPU tensors instead of emulating the kernels
(there is no CPU / eager fallback on the product path; the CPU oracles live in oracle/ and are
test-only).  Runs without a GPU: the refusal happens before any launch."""
import pytest
import torch

from graph_detr4d_b200 import assign, frustum, fused, ops, optim, synthetic as syn
from graph_detr4d_b200.ops import MODE_C, GenLayout, XViewConfig


def test_fused_glue_refuses_cpu():
    with pytest.raises(RuntimeError):
        fused.inverse_sigmoid(torch.rand(4, 3))
    with pytest.raises(RuntimeError):
        fused.ref_update(torch.randn(4, 10), torch.rand(4, 3))
    with pytest.raises(RuntimeError):
        fused.add_layernorm(torch.randn(4, 256), torch.nn.LayerNorm(256))
    with pytest.raises(RuntimeError):
        fused.bias_act_(torch.randn(4, 256), torch.randn(256))


def test_frustum_and_assignment_refuse_cpu():
    with pytest.raises(RuntimeError):
        frustum.frustum_position_input([(4, 6)], syn.make_img_metas(1, 1), 8, 1, syn.PC_RANGE, device="cpu")
    with pytest.raises(RuntimeError):
        assign.BatchedHungarianAssigner3D().assign_layers(torch.randn(1, 1, 8, 10), torch.randn(1, 1, 8, 10),
                                                          [torch.rand(2, 9) + 0.5], [torch.zeros(2, dtype=torch.long)])


def test_optimizer_and_packed_generator_refuse_cpu():
    with pytest.raises(TypeError):
        optim.MultiTensorAdamW([torch.nn.Parameter(torch.zeros(4, 4))])
    cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0)
    vals = [torch.zeros(6, 4, 6, 256)]
    with pytest.raises(RuntimeError):
        ops.xview_forward_gen(cfg, vals, 1, 6, torch.rand(1, 5, 3), torch.zeros(1, 5, 232),
                              GenLayout(cam=224, offsets=128, attn=0, width=232), torch.zeros(1, 6, 4, 4))
