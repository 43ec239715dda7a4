"""Micro-benchmark of pcm_attn_fwd / pcm_attn_bwd on the step's attention shapes (CUDA events).
Usage: python tools/attn_bench.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcm_b200 import ops

dev = torch.device("cuda")
BF = torch.bfloat16
SHAPES = [  # (B, H, Sq, Skv, D)
    (24, 8, 4096, 4096, 40), (8, 8, 4096, 4096, 40), (24, 8, 4096, 77, 40),
    (24, 8, 1024, 1024, 80), (8, 8, 1024, 1024, 80), (24, 8, 256, 256, 160),
]
iters = int(os.environ.get("ITERS", "10"))
for (B, H, Sq, Skv, D) in SHAPES:
    C = H * D
    q = torch.randn(B * Sq, C, device=dev).to(BF)
    k = torch.randn(B * Skv, C, device=dev).to(BF)
    v = torch.randn(B * Skv, C, device=dev).to(BF)
    o = torch.empty_like(q)
    do = torch.randn_like(q)
    lse = torch.empty(B, H, Sq, device=dev, dtype=torch.float32)
    delta = torch.empty_like(lse)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    sc = D ** -0.5

    def timeit(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / iters
    tf = timeit(lambda: ops.attn_fwd(q, k, v, o, lse, B, H, Sq, Skv, D, sc))
    fl = 4.0 * B * H * Sq * Skv * D
    line = f"B={B} H={H} Sq={Sq} Skv={Skv} D={D}: fwd {tf:8.1f} us {fl / tf / 1e6:7.1f} TFLOP/s"
    if B == 8:
        tb = timeit(lambda: ops.attn_bwd(q, k, v, o, do, lse, delta, dq, dk, dv, B, H, Sq, Skv, D, sc))
        line += f" | bwd {tb:8.1f} us {2.5 * fl / tb / 1e6:7.1f} TFLOP/s"
    print(line, flush=True)
