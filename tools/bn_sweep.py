"""Sweep the N tile (block_n) of pcm_gemm on the step's dominant shapes and compare with the
heuristic of ops.pick_block_n (CUDA events, rotating buffers).  Usage: python tools/bn_sweep.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcm_b200 import ops

dev = torch.device("cuda")
BF = torch.bfloat16
SHAPES = [  # (kind, M|(B,H,W), K|Cin, N, residual)
    ("lin", 98304, 384, 320, True), ("lin", 32768, 384, 320, True), ("lin", 98304, 384, 2560, False),
    ("lin", 24576, 704, 640, True), ("lin", 8192, 704, 640, True), ("lin", 6144, 1344, 1280, True),
    ("lin", 2048, 1344, 1280, True), ("lin", 24576, 704, 5120, False), ("lin", 98304, 1344, 320, True),
    ("conv", (24, 64, 64), 320, 320, True), ("conv", (8, 64, 64), 320, 320, True),
    ("conv", (24, 32, 32), 640, 640, True), ("conv", (24, 16, 16), 1280, 1280, True),
    ("conv", (8, 32, 32), 640, 640, True), ("conv", (8, 16, 16), 1280, 1280, True),
]
iters = int(os.environ.get("ITERS", "10"))
for kind, Mx, K, N, res in SHAPES:
    nbuf = 3
    if kind == "lin":
        M = Mx
        xs = [torch.randn(M, K, device=dev).to(BF) for _ in range(nbuf)]
        w = ops.kblock((torch.randn(N, K, device=dev) * K ** -0.5).to(BF))
        prog = [(0, 0, 0, 0, K // 64, 0, 0)]
        Ktot, geo = K, (1, 1)
    else:
        B, H, W = Mx
        M = B * H * W
        xs = [torch.randn(B, H, W, K, device=dev).to(BF) for _ in range(nbuf)]
        w = ops.kblock((torch.randn(N, 9 * K, device=dev) * (9 * K) ** -0.5).to(BF))
        prog = [(0, 0, dw, dh, K // 64, 0, t * K) for t, (dw, dh) in enumerate(ops.TAPS3)]
        Ktot, geo = 9 * K, (W, H)
    outs = [torch.empty(M, N, device=dev, dtype=BF) for _ in range(nbuf)]
    rs = [torch.randn(M, N, device=dev).to(BF) for _ in range(nbuf)] if res else [None] * nbuf
    bias = torch.randn(N, device=dev)
    heur = ops.pick_block_n(M, N)
    row = []
    for bn in (64, 96, 128, 160, 192, 224, 256):
        if bn - 32 >= N:
            continue

        def run(i):
            x = xs[i % nbuf]
            a = [ops.asrc_mat(x)] if kind == "lin" else [ops.asrc_nhwc(x)]
            ops.gemm(a, [ops.bsrc(w)], prog, lin=(kind == "lin"), M=M, N=N, geo=geo, out=outs[i % nbuf], bias=bias,
                     residual=rs[i % nbuf], block_n=bn, ksplit=1)
        for i in range(2):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        row.append((bn, e0.elapsed_time(e1) * 1e3 / iters))
    best = min(row, key=lambda r: r[1])
    hv = dict(row).get(heur)
    print(f"{kind} M={M} K={Ktot} N={N} res={res}: heuristic bn={heur} {hv:.1f} us | best bn={best[0]} {best[1]:.1f} us "
          f"({100 * (hv - best[1]) / hv:.0f} %) | " + " ".join(f"{b}:{t:.1f}" for b, t in row), flush=True)
