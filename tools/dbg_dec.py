import sys, torch, warnings; warnings.filterwarnings("ignore"); sys.path.insert(0,'.')
import graph_detr4d_b200 as g
from graph_detr4d_b200 import synthetic as syn
from tests import helpers as H
from tests.test_decoder_gpu import _build
from oracle.modules_port import build_oracle_attention
for variant, T in [("A",1),("C",2)]:
    sc = H.scene(B=1, T=T, Q=80)
    ref_model = _build(variant, sc["N"], 3, factory=build_oracle_attention)
    model = _build(variant, sc["N"], 3).cuda()
    model.load_state_dict(ref_model.state_dict(), strict=True)
    feats_o = [f.clone().requires_grad_(True) for f in sc["feats"]]
    st_o, r0_o, refs_o = ref_model(feats_o, sc["metas"], 1)
    gout=torch.randn(st_o.shape, generator=torch.Generator().manual_seed(5)); (st_o*gout).sum().backward()
    feats_g = [f.cuda().requires_grad_(True) for f in sc["feats"]]
    g.clear_caches()
    st, r0, refs = model(feats_g, sc["metas"], 1)
    (st*gout.cuda()).sum().backward()
    print(variant, 'st', H.rel_err(st.detach().cpu(), st_o.detach()), 'refs', H.rel_err(refs.detach().cpu(), refs_o.detach()))
    for a,b in zip(feats_g, feats_o):
        d=(a.grad.cpu()-b.grad).abs()
        print('  feat grad rel', H.rel_err(a.grad.cpu(), b.grad), 'max ref', b.grad.abs().max().item(), 'n bad', (d>1e-3*b.grad.abs().max()).sum().item(), 'nnz', (b.grad!=0).sum().item(), (a.grad!=0).sum().item())
