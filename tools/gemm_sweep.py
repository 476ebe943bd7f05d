import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safevla_b200 import ops, _lib as L
dev = torch.device("cuda:0"); bf = torch.bfloat16
def bench(name, M, N, K, ta=False, tb=True, bias=False, relu=False, res=False, out_f32=False, acc=False, mask=False):
    A = torch.randn((K, M) if ta else (M, K), device=dev, dtype=bf)
    B = torch.randn((N, K) if tb else (K, N), device=dev, dtype=bf)
    C = torch.zeros(M, N, device=dev, dtype=torch.float32 if out_f32 else bf)
    b = torch.zeros(N, device=dev) if bias else None
    r = torch.randn(M, N, device=dev, dtype=C.dtype) if res else None
    aux = torch.randn(M, N, device=dev, dtype=C.dtype) if mask else None
    f = lambda: ops.gemm(A, B, C, trans_a=ta, trans_b=tb, bias=b, residual=r, aux=aux, epilogue=L.EPI_RELU_MASK if mask else (L.EPI_RELU if relu else L.EPI_NONE), accumulate=acc, impl=2)
    for _ in range(3): f()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
    e[0].record()
    for i in range(8):
        f(); e[i+1].record()
    torch.cuda.synchronize()
    t = min(e[i].elapsed_time(e[i+1]) for i in range(8)) * 1e-3
    print(f"{name:34s} M={M:7d} N={N:5d} K={K:6d}  {t*1e6:8.1f} us  {2*M*N*K/t/1e12:7.1f} TF/s")
M = 119808
bench("fwd K512 N2048 plain", M, 2048, 512)
bench("fwd K512 N2048 bias+relu", M, 2048, 512, bias=True, relu=True)
bench("fwd K512 N512 plain", M, 512, 512)
bench("fwd K512 N512 bias+res", M, 512, 512, bias=True, res=True)
bench("fwd K512 N1536 bias (qkv)", M, 1536, 512, bias=True)
bench("fwd K2048 N512 bias+res (ffn2)", M, 512, 2048, bias=True, res=True)
bench("fwd K2048 N2048 plain", M, 2048, 2048)
bench("fwd K384 N512 relu (compressor)", 86016, 512, 384, bias=True, relu=True)
bench("dgrad K2048 N512 (B MN)", M, 512, 2048, tb=False)
bench("dgrad K512 N2048 (B MN)", M, 2048, 512, tb=False)
bench("dgrad K512 N512 (B MN)", M, 512, 512, tb=False)
bench("dgrad K512 N2048 relu-mask", M, 2048, 512, tb=False, mask=True)
bench("dgrad K512 N512 relu-mask", M, 512, 512, tb=False, mask=True)
bench("dgrad K2048 N512 +res", M, 512, 2048, tb=False, res=True)
bench("dgrad K1536 N512 +res", M, 512, 1536, tb=False, res=True)
bench("dgrad K1536 N512 (B MN)", M, 512, 1536, tb=False)
bench("wgrad 2048x512 (MN,MN) f32 acc", 2048, 512, M, ta=True, tb=False, out_f32=True, acc=True)
bench("wgrad 512x2048 f32 acc", 512, 2048, M, ta=True, tb=False, out_f32=True, acc=True)
bench("wgrad 512x512 f32 acc", 512, 512, M, ta=True, tb=False, out_f32=True, acc=True)
bench("wgrad 1536x512 f32 acc", 1536, 512, M, ta=True, tb=False, out_f32=True, acc=True)
bench("small M=1024 N2048 K512", 1024, 2048, 512, bias=True, relu=True)
bench("small M=8192 N1536 K512", 8192, 1536, 512)
bench("small M=8192 N512 K1536", 8192, 512, 1536, out_f32=True, res=True)
