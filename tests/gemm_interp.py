"""Test-side interpreter of `pcm_gemm` descriptors for Linear-type launches (lin = 1): what the CUDA
kernel computes, restated with torch on the raw operand pointers - K program entries (a_src, b_src,
dw, dh, nchunks, a_c0, b_k0 [, n_lo, n_hi]) as documented in include/pcm_b200.h, TMA zero fill past an
A source's rows, K-blocked B sources, fp32 accumulation, epilogue  act(acc + bias + residual) -> bf16.

Used by the CPU tests of the HOST launch plans (which layers are stacked, which LoRA block feeds which
columns / rows).  Test infrastructure only: nothing under pcm_b200/ imports it and the product has no CPU
path."""
import ctypes

import torch

BF16 = torch.bfloat16


def mat(ptr, rows, cols, ld):
    """bf16 [rows, cols] view (row stride ld elements) of raw memory."""
    n = (rows - 1) * ld + cols
    buf = (ctypes.c_uint16 * n).from_address(ptr)
    return torch.frombuffer(buf, dtype=torch.int16).view(BF16).as_strided((rows, cols), (ld, 1))


def b_matrix(b):
    """[N, K] view of a pcm_bsrc (row-major or K-blocked [K/64][N][64])."""
    if b.kblocked:
        return mat(b.ptr, (b.K // 64) * b.N, 64, 64).view(b.K // 64, b.N, 64).permute(1, 0, 2).reshape(b.N, b.K)
    return mat(b.ptr, b.N, b.K, b.ld)


def interp_gemm(a_srcs, b_srcs, prog, *, lin, M, N, out, bias=None, residual=None, act=0, rowvec=None,
                alpha=1.0, **kw):
    assert lin, "Linear-type launches only"
    assert rowvec is None and alpha == 1.0
    acc = torch.zeros(M, N, dtype=torch.float32)
    for e in prog:
        a, b = a_srcs[e[0]], b_srcs[e[1]]
        assert e[2] == 0 and e[3] == 0
        kk = 64 * e[4]
        rows = min(M, a.W)
        A = mat(a.ptr, rows, a.C, a.sW)[:, e[5]:e[5] + kk].float()
        lo, hi = (e[7], e[8]) if (len(e) > 7 and e[8]) else (0, N)
        Bm = b_matrix(b)[lo:hi, e[6]:e[6] + kk].float()
        acc[:rows, lo:hi] += A @ Bm.t()
    if bias is not None:
        acc += bias[:N].float()
    if residual is not None:
        acc += residual.float()
    if act == 1:
        acc = torch.nn.functional.silu(acc)
    else:
        assert act == 0
    assert out.shape == (M, N)
    out.copy_(acc.to(out.dtype))
    return out


def interp_wgrad(p_src, q_src, out, *, lin, M, os_row, os_col, alpha=1.0, q_c0=0, taps=((0, 0),),
                 tap_off=(0,), **kw):
    """out[tap_off + ch*os_row + r*os_col] += alpha * sum_m P[m, ch] * Q[m, q_c0 + r], r < 64 (pcm_wgrad, lin)."""
    assert lin and len(taps) == 1 and out.dtype == torch.float32
    rows = min(M, p_src.W, q_src.W)
    P = mat(p_src.ptr, rows, p_src.C, p_src.sW).float()
    Q = mat(q_src.ptr, rows, q_src.C, q_src.sW)[:, q_c0:q_c0 + 64].float()
    o = out.as_strided((P.shape[1], 64), (os_row, os_col), out.storage_offset() + tap_off[0])
    o += alpha * (P.t() @ Q)
    return out


def build_net(cfg, seed=3, lora_b_std=0.2):
    """UNetB200 on CPU in record-only mode, with the operand copies `pcm_lora_refresh` would write
    (bf16 A and s*B through the per-layer views)."""
    from pcm_b200 import ops, weights
    from pcm_b200.unet import UNetB200
    sd = weights.synthetic_state_dict(cfg, seed, lora_b_std=lora_b_std)
    old = ops.DRY_RUN
    ops.DRY_RUN = []
    try:
        net = UNetB200(cfg, sd, "cpu", lora=True, need_backward=True)
    finally:
        ops.DRY_RUN = old
    for L in net.lora_layers:
        lo = L.lora
        na, nb = lo.a_fwd.numel(), lo.sb_fwd.numel()
        lo.a_fwd.copy_(net.lora_master[lo.a_off:lo.a_off + na].view_as(lo.a_fwd).to(BF16))
        sB = (net.scale * net.lora_master[lo.b_off:lo.b_off + nb].view_as(lo.sb_fwd)).to(BF16)
        lo.sb_fwd.copy_(sB)
        lo.sb_t.copy_(sB.t())
        if L.kind != "conv" or L.k == 1:      # the tap-major transposed copy of 3x3 adapters is not needed here
            lo.a_t.copy_(lo.a_fwd.t())
    return net, sd


def lora_linear_ref(sd, name, x, scale, lora_rows=None, bias=True):
    """peft LoRA Linear on bf16 operands: x W^T (+ b) + s (x A^T) B^T, the adapter on the leading rows."""
    x = x.float()
    W = sd[name + ".weight"]
    W = W.reshape(W.shape[0], -1)                       # nn.Linear or 1x1 Conv2d
    y = x @ W.to(BF16).float().t()
    if bias and (name + ".bias") in sd:
        y = y + sd[name + ".bias"].float()
    if (name + ".lora_A.weight") in sd:
        rows = x.shape[0] if lora_rows is None else lora_rows
        A = sd[name + ".lora_A.weight"]
        A = A.reshape(A.shape[0], -1).to(BF16).float()
        sB = scale * sd[name + ".lora_B.weight"]
        sB = sB.reshape(sB.shape[0], -1).to(BF16).float()
        T = (x[:rows] @ A.t()).to(BF16).float()
        y[:rows] += T @ sB.t()
    return y
