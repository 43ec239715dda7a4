"""Micro-benchmark of the HBM-bound normalisation / activation kernels on the step's shapes
(CUDA events, rotating buffers so every iteration misses L2).  Usage: python tools/norm_bench.py [gn|ln|geglu]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcm_b200 import ops

dev = torch.device("cuda")
BF = torch.bfloat16
iters = int(os.environ.get("ITERS", "10"))
which = sys.argv[1:] or ["gn", "ln", "geglu"]
NB = 3


def timeit(fn):
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


if "gn" in which:
    for (B, HW, C, silu) in [(24, 4096, 320, True), (24, 1024, 640, True), (24, 256, 1280, True), (8, 4096, 320, True),
                             (24, 4096, 320, False)]:
        xs = [torch.randn(B, HW, C, device=dev).to(BF) for _ in range(NB)]
        dys = [torch.randn(B, HW, C, device=dev).to(BF) for _ in range(NB)]
        outs = [torch.empty(B, HW, C, device=dev, dtype=BF) for _ in range(NB)]
        gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        stats = torch.empty(B, 32, 2, device=dev)
        red = torch.empty(B, 32, 2, device=dev)
        tf = timeit(lambda i: ops.groupnorm_fwd(xs[i % NB], None, gamma, beta, 1e-5, silu, outs[i % NB], stats, B, HW))
        tb = timeit(lambda i: ops.groupnorm_bwd(dys[i % NB], xs[i % NB], None, gamma, beta, 1e-5, silu, stats, red, None,
                                                outs[i % NB], None, B, HW))
        n = B * HW * C * 2
        print(f"GN B={B} HW={HW} C={C} silu={silu}: fwd {tf:7.1f} us ({3 * n / tf / 1e3:6.0f} GB/s alg.)  "
              f"bwd {tb:7.1f} us ({5 * n / tb / 1e3:6.0f} GB/s alg.)", flush=True)
if "ln" in which:
    for (M, C) in [(98304, 320), (24576, 640), (6144, 1280), (32768, 320)]:
        xs = [torch.randn(M, C, device=dev).to(BF) for _ in range(NB)]
        dys = [torch.randn(M, C, device=dev).to(BF) for _ in range(NB)]
        outs = [torch.empty(M, C, device=dev, dtype=BF) for _ in range(NB)]
        gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        stats = torch.empty(M, 2, device=dev)
        tf = timeit(lambda i: ops.layernorm_fwd(xs[i % NB], gamma, beta, outs[i % NB], stats))
        tb = timeit(lambda i: ops.layernorm_bwd(dys[i % NB], xs[i % NB], gamma, stats, None, outs[i % NB]))
        n = M * C * 2
        print(f"LN M={M} C={C}: fwd {tf:7.1f} us ({2 * n / tf / 1e3:6.0f} GB/s alg.)  bwd {tb:7.1f} us "
              f"({3 * n / tb / 1e3:6.0f} GB/s alg.)", flush=True)
if "geglu" in which:
    for (M, F) in [(98304, 1280), (24576, 2560), (32768, 1280)]:
        us = [torch.randn(M, 2 * F, device=dev).to(BF) for _ in range(NB)]
        outs = [torch.empty(M, F, device=dev, dtype=BF) for _ in range(NB)]
        tf = timeit(lambda i: ops.geglu_fwd(us[i % NB], outs[i % NB]))
        print(f"GEGLU M={M} F={F}: fwd {tf:7.1f} us ({M * F * 6 / tf / 1e3:6.0f} GB/s alg.)", flush=True)
