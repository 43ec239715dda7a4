"""Debug helper: run one split-K convolution with a NaN-prefilled workspace and report unwritten
workspace elements / wrong slices (GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from pcm_b200 import ops

BF = torch.bfloat16
dev = torch.device("cuda:0")
B, H, W, Cin, Cout = [int(a) for a in sys.argv[1:6]] if len(sys.argv) > 5 else (8, 8, 8, 1280, 64)
g = torch.Generator().manual_seed(1)
x = torch.randn(B, H, W, Cin, generator=g).to(dev).to(BF)
w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (9 * Cin) ** -0.5).to(dev).to(BF)
wm = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
M = B * H * W
prog = [(0, 0, dw, dh, Cin // 64, 0, t * Cin) for t, (dw, dh) in enumerate(ops.TAPS3)]
bn, ks = ops.pick_tiling(M, Cout, len(prog) * (Cin // 64))
print("M", M, "N", Cout, "nkb", len(prog) * Cin // 64, "bn", bn, "ks", ks)
for trial in range(3):
    ws = torch.full((ks, M, Cout), float("nan"), device=dev)
    out = torch.empty(M, Cout, device=dev, dtype=BF)
    ops.gemm([ops.asrc_nhwc(x)], [ops.bsrc(wm)], prog, lin=False, M=M, N=Cout, geo=(W, H), out=out,
             block_n=bn, ksplit=ks, splitk_ws=ws)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, padding=1).permute(0, 2, 3, 1).reshape(M, Cout)
    nan = torch.isnan(ws)
    used = [(~nan[s]).any().item() for s in range(ks)]
    print(f"trial {trial}: slices touched {sum(used)} / {ks}; NaN inside touched slices: "
          f"{sum(nan[s].sum().item() for s in range(ks) if used[s])}; out max err "
          f"{(out.float() - ref).abs().max().item():.4f} (ref max {ref.abs().max().item():.3f})")
    tot = torch.nan_to_num(ws, nan=0.0).sum(0)
    print("   sum of slices vs ref max err", (tot - ref).abs().max().item())
    bad = ((out.float() - ref).abs() > 0.05 * ref.abs().max()).nonzero()
    if len(bad):
        print("   bad elements (m, n):", bad[:10].tolist(), "count", len(bad))
        for s in range(ks):
            if used[s] and nan[s].any():
                idx = nan[s].nonzero()
                print(f"   slice {s}: {len(idx)} NaN, first {idx[:4].tolist()}")

# ---- the failing test's configuration: bias + per-image row vector + residual + SiLU ----------
print("with bias / rowvec / residual / act=1")
bias = torch.randn(Cout, device=dev)
res = torch.randn(B, H, W, Cout, generator=g).to(dev).to(BF)
rowvec = torch.randn(B, Cout, generator=g).to(dev).to(BF)
for trial in range(3):
    ws = torch.full((ks, M, Cout), float("nan"), device=dev)
    out = torch.empty(B, H, W, Cout, device=dev, dtype=BF)
    ops.gemm([ops.asrc_nhwc(x)], [ops.bsrc(wm)], prog, lin=False, M=M, N=Cout, geo=(W, H), out=out.view(-1, Cout),
             bias=bias, rowvec=rowvec, residual=res.view(-1, Cout), act=1, splitk_ws=ws if trial < 2 else None)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1)
    ref = F.silu(ref + rowvec.float()[:, :, None, None] + res.float().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    bad = (err > 0.05 * ref.abs().max()).nonzero()
    print(f"trial {trial} (own ws: {trial < 2}): max err {err.max().item():.4g}; bad {len(bad)}; first {bad[:6].tolist()}")
    if trial < 2:
        nan = torch.isnan(ws)
        print("   NaN in first 20 slices:", nan[:20].sum().item(), " in last 2:", nan[20:].sum().item())
