"""One call of each fp32 GEMM implementation at the decoder's Linear shape (for ncu).  (dev tool)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graph_detr4d_b200 import gemm as G
torch.backends.cuda.matmul.allow_tf32 = False
M, N, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (900, 256, 256)))
a, b, bias = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), torch.randn(N, device="cuda")
o = torch.empty(M, N, device="cuda")
for _ in range(3):
    G.gemm(a, b, bias, out=o, impl="simt")
    torch.addmm(bias, a, b.t(), out=o)
torch.cuda.synchronize()
