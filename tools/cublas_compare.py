#!/usr/bin/env python
"""The fusion block's GEMM shapes through this library (tcgen05 pair kernel, fused epilogues) and through cuBLASLt
(torch.nn.functional.linear / torch.mm / torch.addmm in bf16) on the SAME box, interleaved, mean of 20 launches after
warm-up.  Context for roofline.frac: MEASURED_PEAKS' 1383 TF/s is cuBLAS on a large square problem; the update's
GEMMs have K = 512 .. 2048 on one side and write [M, N] bf16 tiles whose epilogues (bias, ReLU, residual) cuBLAS
either fuses (bias) or leaves to separate kernels (ReLU, residual: timed here as the extra torch kernels they need).

    python tools/cublas_compare.py [rows]          # rows of the [rows, *] activations (default 479232 = 4096 x 117)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from safevla_b200 import _lib as L  # noqa: E402
from safevla_b200 import ops  # noqa: E402

dev, bf = torch.device("cuda:0"), torch.bfloat16
M = int(sys.argv[1]) if len(sys.argv) > 1 else 479232


def t_of(fn, n=20):
    for _ in range(4):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def row(name, flops, ours, lib, lib_note):
    # interleave twice so both see the same thermal / power state
    to = min(t_of(ours), t_of(ours))
    tl = min(t_of(lib), t_of(lib))
    print(f"{name:44s} ours {to*1e6:8.1f} us {flops/to/1e12:7.1f} TF/s | cuBLASLt {tl*1e6:8.1f} us {flops/tl/1e12:7.1f} TF/s"
          f"  ({lib_note})  ratio {tl/to:5.2f}x", flush=True)


def main():
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g).to(bf)  # noqa: E731
    for N, K, kind in ((2048, 512, "relu"), (1536, 512, "bias"), (512, 512, "res"), (512, 2048, "res")):
        x, w, b = rn(M, K), rn(N, K) * K ** -0.5, torch.randn(N, device=dev, generator=g)
        bb = b.to(bf)
        out = torch.empty(M, N, device=dev, dtype=bf)
        res = rn(M, N) if kind == "res" else None
        fl = 2.0 * M * N * K
        if kind == "relu":
            bits = torch.empty(M, N // 32, device=dev, dtype=torch.int32)
            row(f"fwd [{M},{K}]x[{N},{K}]^T bias+ReLU", fl,
                lambda: ops.gemm(x, w, out, trans_b=True, bias=b, epilogue=L.EPI_RELU),
                lambda: torch.relu_(F.linear(x, w, bb)), "linear + relu_")
            row(f"fwd [{M},{K}]x[{N},{K}]^T bias+ReLU+bit record", fl,
                lambda: ops.gemm(x, w, out, trans_b=True, bias=b, epilogue=L.EPI_RELU_BITS, aux=bits),
                lambda: torch.relu_(F.linear(x, w, bb)), "linear + relu_")
        elif kind == "bias":
            row(f"fwd [{M},{K}]x[{N},{K}]^T bias", fl,
                lambda: ops.gemm(x, w, out, trans_b=True, bias=b),
                lambda: F.linear(x, w, bb, ), "linear")
        else:
            row(f"fwd [{M},{K}]x[{N},{K}]^T bias+residual", fl,
                lambda: ops.gemm(x, w, out, trans_b=True, bias=b, residual=res),
                lambda: F.linear(x, w, bb).add_(res), "linear + add_")
        row(f"fwd [{M},{K}]x[{N},{K}]^T matmul only", fl,
            lambda: ops.gemm(x, w, out, trans_b=True),
            lambda: torch.mm(x, w.t(), out=out), "mm")
    # data gradients  dx = dy W  (B row-major [K, N])
    for N, K in ((512, 2048), (2048, 512), (512, 1536), (512, 512)):
        dy, w = rn(M, K), rn(K, N) * K ** -0.5
        out = torch.empty(M, N, device=dev, dtype=bf)
        row(f"dgrad [{M},{K}]x[{K},{N}]", 2.0 * M * N * K,
            lambda: ops.gemm(dy, w, out, trans_b=False), lambda: torch.mm(dy, w, out=out), "mm")
    # weight gradients  dW += dy^T x  (fp32 accumulate into the gradient arena; cuBLAS: bf16 out, no accumulate)
    for No, Ki in ((2048, 512), (512, 2048), (1536, 512), (512, 512)):
        dy, x = rn(M, No), rn(M, Ki)
        gw, gb = torch.zeros(No, Ki, device=dev), torch.zeros(No, device=dev)
        outb = torch.empty(No, Ki, device=dev, dtype=bf)
        row(f"wgrad [{No},{M}]x[{M},{Ki}] (+bias grad, fp32 +=)", 2.0 * M * No * Ki,
            lambda: ops.gemm(dy, x, gw, trans_a=True, trans_b=False, accumulate=True, colsum_a=gb),
            lambda: (torch.mm(dy.t(), x, out=outb), dy.sum(0)), "mm bf16 out + sum(0)")


if __name__ == "__main__":
    main()
