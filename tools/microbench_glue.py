import torch, time
import torch.nn.functional as F
dev='cuda'
M,K,N=900,256,256
x=torch.randn(M,K,device=dev); g=torch.randn(M,N,device=dev); W=torch.randn(N,K,device=dev); b=torch.randn(N,device=dev)
ones=torch.ones(1,M,device=dev)
def t(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    # capture in graph to exclude launch overhead
    gr=torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20): fn()
    gr.replay(); torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): gr.replay()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e)/200*1e3
print('addmm fwd        us', t(lambda: torch.addmm(b,x,W.t())))
print('dW g.t()@x       us', t(lambda: g.t()@x))
print('dW (x.t()@g).t() us', t(lambda: (x.t()@g).t()))
print('dX g@W           us', t(lambda: g@W))
print('db sum(0)        us', t(lambda: g.sum(0)))
print('db ones@g        us', t(lambda: ones@g))
for N2 in (6,96,128,512,10):
    g2=torch.randn(M,N2,device=dev); W2=torch.randn(N2,K,device=dev)
    print(f'N={N2}: dW g.t()@x', t(lambda: g2.t()@x), ' (x.t()@g).t()', t(lambda: (x.t()@g2).t()), ' dX', t(lambda: g2@W2), ' db sum', t(lambda: g2.sum(0)), ' ones@g', t(lambda: ones@g2))
q=torch.randn(1,8,900,32,device=dev,requires_grad=True); k=torch.randn(1,8,900,32,device=dev,requires_grad=True); v=torch.randn(1,8,900,32,device=dev,requires_grad=True)
go=torch.randn(1,8,900,32,device=dev)
from torch.nn.attention import sdpa_kernel, SDPBackend
def sd(backend):
    def f():
        with sdpa_kernel(backend):
            o=F.scaled_dot_product_attention(q,k,v)
        o.backward(go)
        q.grad=None;k.grad=None;v.grad=None
    return f
for be in (SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH):
    try: print('sdpa fwd+bwd', be, t(sd(be)))
    except Exception as ex: print('sdpa', be, 'failed', str(ex)[:100])
def manual():
    s=torch.matmul(q,k.transpose(-1,-2))*(32**-0.5)
    p=s.softmax(-1); o=torch.matmul(p,v); o.backward(go)
    q.grad=None;k.grad=None;v.grad=None
print('manual attn fwd+bwd', t(manual))
