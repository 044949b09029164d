"""Accuracy + timing table of the tcgen05 3xTF32 GEMM against fp64 and against cuBLAS fp32 (dev tool)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graph_detr4d_b200 import gemm as G

torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
gen = torch.Generator(device="cpu").manual_seed(0)
rows = []

def check(name, M, N, K, a_t=False, b_t=False, batch=0, bias=False, relu=False, time=True):
    sh = lambda r, c: ((batch, r, c) if batch else (r, c))
    a = torch.randn(sh(K, M) if a_t else sh(M, K), generator=gen).to(dev)
    b = torch.randn(sh(K, N) if b_t else sh(N, K), generator=gen).to(dev)
    bi = torch.randn(N, generator=gen).to(dev) if bias else None
    A = a.transpose(-1, -2) if a_t else a
    B = b if b_t else b.transpose(-1, -2)
    ref64 = A.double() @ B.double()
    if bias: ref64 = ref64 + bi.double()
    if relu: ref64 = ref64.relu()
    lib32 = A @ B
    if bias: lib32 = lib32 + bi
    if relu: lib32 = lib32.relu()
    out = G.gemm(a, b, bi, relu, a_t, b_t, impl="tf32x3")
    out_s = G.gemm(a, b, bi, relu, a_t, b_t, impl="simt")
    torch.cuda.synchronize()
    den = float(ref64.abs().max())
    e_ours = float((out.double() - ref64).abs().max()) / den
    e_simt = float((out_s.double() - ref64).abs().max()) / den
    e_lib = float((lib32.double() - ref64).abs().max()) / den
    t_ours = t_lib = t_simt = None
    if time:
        def tm(fn):
            # device time: 20 calls captured in a CUDA graph (the python / ctypes host side of a call is
            # longer than these kernels), replayed 10 times
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                for _ in range(3): fn()
                side.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    for _ in range(20): fn()
                graph.replay(); side.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(side)
                for _ in range(10): graph.replay()
                e.record(side); side.synchronize()
            return s.elapsed_time(e) / 200 * 1e3
        o = torch.empty_like(out)
        t_ours = tm(lambda: G.gemm(a, b, bi, relu, a_t, b_t, out=o, impl="tf32x3"))
        t_simt = tm(lambda: G.gemm(a, b, bi, relu, a_t, b_t, out=o, impl="simt"))
        t_lib = tm(lambda: torch.matmul(A, B, out=o))
    r = dict(name=name, M=M, N=N, K=K, a_t=a_t, b_t=b_t, batch=batch, err_ours=e_ours, err_simt=e_simt, err_cublas=e_lib, us_ours=t_ours, us_simt=t_simt, us_cublas=t_lib)
    rows.append(r)
    print(json.dumps(r), flush=True)

check("tiny_NT", 128, 32, 32, time=False)
check("tiny_NT_k64", 128, 64, 64, time=False)
check("fwd_256", 900, 256, 256, bias=True)
check("fwd_relu_512", 900, 512, 256, bias=True, relu=True)
check("fwd_k512", 900, 256, 512)
check("gen_232", 900, 232, 256, bias=True)
check("reg_10", 900, 10, 256, bias=True)
check("ktail_40", 70, 36, 40, time=False)
check("dgrad_NN", 900, 256, 256, b_t=True)
check("dgrad_NN_512", 900, 256, 512, b_t=True)
check("wgrad_TN", 256, 256, 900, a_t=True, b_t=True, batch=24)
check("wgrad_TN_512", 512, 256, 900, a_t=True, b_t=True, batch=6)
check("wide_vproj", 900, 32, 256, batch=8)
check("qk", 900, 900, 32, batch=8)
check("pv", 900, 32, 900, b_t=True, batch=8)
check("bigM_7200", 7200, 256, 256, bias=True)
check("bigM_7200_k512", 7200, 512, 512)
json.dump(rows, open("gpurun_out/r2_gemm_check.json", "w"), indent=1)
