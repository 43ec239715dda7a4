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


def a_nhwc(a):
    """bf16 [B, H, W, C] strided view of a pcm_asrc."""
    n = (a.B - 1) * a.sB + (a.H - 1) * a.sH + (a.W - 1) * a.sW + a.C
    buf = (ctypes.c_uint16 * n).from_address(a.ptr)
    return torch.frombuffer(buf, dtype=torch.int16).view(BF16).as_strided((a.B, a.H, a.W, a.C), (a.sB, a.sH, a.sW, 1))


def shifted_rows(a, dw, dh, Wo, Ho, Bo):
    """[Bo*Ho*Wo, C] fp32: row (b, h, w) = source pixel (b, h + dh, w + dw), zero outside the source
    (what the 4-D TMA box load of the implicit GEMM delivers)."""
    x = a_nhwc(a).float()
    out = torch.zeros(Bo, Ho, Wo, a.C)
    b1 = min(Bo, a.B)
    h0, h1 = max(0, -dh), min(Ho, a.H - dh)
    w0, w1 = max(0, -dw), min(Wo, a.W - dw)
    if h1 > h0 and w1 > w0:
        out[:b1, h0:h1, w0:w1] = x[:b1, h0 + dh:h1 + dh, w0 + dw:w1 + dw]
    return out.reshape(Bo * Ho * Wo, a.C)


def interp_gemm(a_srcs, b_srcs, prog, *, lin, M, N, out, geo=(1, 1), bias=None, residual=None, act=0,
                rowvec=None, alpha=1.0, round_bf16=False, **kw):
    """Linear (lin) and implicit-convolution launches.  `out` is the destination view itself (dense
    [M, N] or a strided NHWC plane), so out_strides / epi need no separate handling."""
    Wo, Ho = geo
    Bo = M // (Wo * Ho)
    acc = torch.zeros(M, N, dtype=torch.float32)
    for e in prog:
        a, b = a_srcs[e[0]], b_srcs[e[1]]
        kk = 64 * e[4]
        lo, hi = (e[7], e[8]) if (len(e) > 7 and e[8]) else (0, N)
        Bm = b_matrix(b)[lo:hi, e[6]:e[6] + kk].float()
        if lin:
            assert e[2] == 0 and e[3] == 0
            rows = min(M, a.W)
            A = mat(a.ptr, rows, a.C, a.sW)[:, e[5]:e[5] + kk].float()
            acc[:rows, lo:hi] += A @ Bm.t()
        else:
            A = shifted_rows(a, e[2], e[3], Wo, Ho, Bo)[:, e[5]:e[5] + kk]
            acc[:, lo:hi] += A @ Bm.t()
    acc = acc * alpha
    if bias is not None:
        acc += bias[:N].float()
    if rowvec is not None:                       # one row vector per sample (time embedding)
        acc = (acc.view(Bo, Ho * Wo, N) + rowvec[:, :N].float().unsqueeze(1)).reshape(M, N)
    if residual is not None:
        acc += residual.float().reshape(M, N)
    if act == 1:
        acc = torch.nn.functional.silu(acc)
    else:
        assert act == 0
    if out.dtype == torch.float32 and round_bf16:
        acc = acc.to(BF16).float()
    assert out.numel() == M * N
    out.copy_(acc.view(out.shape).to(out.dtype))
    return out


def interp_wgrad(p_src, q_src, out, *, lin, M, os_row, os_col, alpha=1.0, q_c0=0, taps=((0, 0),),
                 tap_off=(0,), **kw):
    """out[tap_off + ch*os_row + r*os_col] += alpha * sum_m P[m, ch] * Q[m, q_c0 + r], r < 64 (pcm_wgrad, lin)."""
    assert out.dtype == torch.float32
    if lin:
        assert len(taps) == 1
        rows = min(M, p_src.W, q_src.W)
        P = mat(p_src.ptr, rows, p_src.C, p_src.sW).float()
        Q = mat(q_src.ptr, rows, q_src.C, q_src.sW)[:, q_c0:q_c0 + 64].float()
        o = out.as_strided((P.shape[1], 64), (os_row, os_col), out.storage_offset() + tap_off[0])
        o += alpha * (P.t() @ Q)
        return out
    Wo, Ho = kw["geo"]
    Bo = M // (Wo * Ho)
    Q = a_nhwc(q_src).float().reshape(-1, q_src.C)[:M, q_c0:q_c0 + 64]
    for (dw, dh), off in zip(taps, tap_off):
        P = shifted_rows(p_src, dw, dh, Wo, Ho, Bo)
        o = out.as_strided((P.shape[1], 64), (os_row, os_col), out.storage_offset() + off)
        o += alpha * (P.t() @ Q)
    return out


def build_net(cfg, seed=3, lora_b_std=0.2, sd=None):
    """UNetB200 on CPU in record-only mode, with the operand copies `pcm_lora_refresh` would write
    (bf16 A and s*B through the per-layer views)."""
    from pcm_b200 import ops, weights
    from pcm_b200.unet import UNetB200
    if sd is None:
        sd = weights.synthetic_state_dict(cfg, seed, lora_b_std=lora_b_std)
    old = ops.DRY_RUN
    ops.DRY_RUN = []
    try:
        net = UNetB200(cfg, sd, "cpu", lora=True, need_backward=True)
    finally:
        ops.DRY_RUN = old
    refresh_operands(net)
    return net, sd


def refresh_operands(net):
    """What `pcm_lora_refresh` writes: the bf16 operand copies A, s*B, (s*B)^T and the per-tap A^T of every
    LoRA layer, from the fp32 master buffer, through the per-layer views."""
    for L in net.lora_layers:
        lo = L.lora
        na, nb = lo.a_fwd.numel(), lo.sb_fwd.numel()
        lo.a_fwd.copy_(net.lora_master[lo.a_off:lo.a_off + na].view_as(lo.a_fwd).to(BF16))
        sB = (net.scale * net.lora_master[lo.b_off:lo.b_off + nb].view_as(lo.sb_fwd)).to(BF16)
        lo.sb_fwd.copy_(sB)
        lo.sb_t.copy_(sB.t())
        taps = L.k * L.k if L.kind == "conv" else 1
        # A^T per tap: a_t[c, t*r + j] = a_fwd[j, t*cin + c]  (the dgrad K program reads column block t)
        lo.a_t.copy_(lo.a_fwd.view(net.r, taps, L.cin).permute(2, 1, 0).reshape(L.cin, taps * net.r))


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
