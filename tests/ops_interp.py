"""Test-side torch semantics of every `pcm_b200.ops` wrapper the UNet sequencing code calls, so that the
HOST logic of pcm_b200/unet.py (tape, backward walk, merged student + teacher pass, grouped layers,
gradient buffer layout) can be executed and compared with the oracle on CPU.

`install(monkeypatch)` replaces the wrappers in `pcm_b200.ops` for the duration of one test.  The
semantics follow include/pcm_b200.h (the same statements the GPU tests check the CUDA kernels against,
tests/test_ops_gpu.py / test_gemm_gpu.py / test_attn_gpu.py).  Test infrastructure only: nothing under
pcm_b200/ imports this file, and the product raises when libpcm_b200.so is missing."""
import math

import torch

from gemm_interp import BF16, interp_gemm, interp_wgrad

F = torch.nn.functional


def _gn(x, B, HW, G, gamma, beta, eps, silu):
    C = x.shape[-1]
    xg = x.view(B, HW, G, C // G)
    mean = xg.mean(dim=(1, 3), keepdim=True)
    var = xg.var(dim=(1, 3), unbiased=False, keepdim=True)
    rstd = (var + eps).rsqrt()
    y = ((xg - mean) * rstd).view(B * HW, C) * gamma + beta
    return (F.silu(y) if silu else y), mean.view(B, G), rstd.view(B, G)


def _cat(x1, x2):
    return x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], -1)


def groupnorm_fwd(x1, x2, gamma, beta, eps, silu, out, stats, B, HW, G=32):
    y, mean, rstd = _gn(_cat(x1, x2), B, HW, G, gamma, beta, eps, silu)
    out.copy_(y.to(out.dtype))
    stats.copy_(torch.stack([mean, rstd], -1))
    return out


def groupnorm_bwd(dy, x1, x2, gamma, beta, eps, silu, stats, red, add, dx1, dx2, B, HW, G=32, colsum=None):
    x = _cat(x1, x2).requires_grad_(True)
    with torch.enable_grad():
        y, _, _ = _gn(x, B, HW, G, gamma, beta, eps, silu)
        (y * dy.float()).sum().backward()
    dx = x.grad
    if add is not None:
        dx = dx + add.float().reshape(dx.shape)
    if colsum is not None:
        colsum.copy_(dx.view(B, HW, -1).sum(1))
    C1 = x1.shape[-1]
    dx1.copy_(dx[:, :C1].to(dx1.dtype))
    if dx2 is not None:
        dx2.copy_(dx[:, C1:].to(dx2.dtype))


def layernorm_fwd(x, gamma, beta, out, stats, eps=1e-5):
    xf = x.float()
    mean = xf.mean(-1, keepdim=True)
    rstd = (xf.var(-1, unbiased=False, keepdim=True) + eps).rsqrt()
    out.copy_(((xf - mean) * rstd * gamma + beta).to(out.dtype))
    stats.copy_(torch.cat([mean, rstd], -1))
    return out


def layernorm_bwd(dy, x, gamma, stats, add, dx):
    xf = x.float().requires_grad_(True)
    with torch.enable_grad():
        y = F.layer_norm(xf, (xf.shape[-1],), gamma, None, 1e-5)
        (y * dy.float()).sum().backward()
    g = xf.grad if add is None else xf.grad + add.float()
    dx.copy_(g.to(dx.dtype))
    return dx


def _heads(t, B, S, H, D):
    return t[:, :H * D].float().reshape(B, S, H, D).permute(0, 2, 1, 3)


def attn_fwd(q, k, v, out, lse, B, H, Sq, Skv, D, scale):
    for b in range(B):                      # one sample at a time: [H, Sq, Skv] scores stay small
        rq, rk = slice(b * Sq, (b + 1) * Sq), slice(b * Skv, (b + 1) * Skv)
        Q, K, V = _heads(q[rq], 1, Sq, H, D), _heads(k[rk], 1, Skv, H, D), _heads(v[rk], 1, Skv, H, D)
        s = (Q @ K.transpose(-1, -2)) * scale
        lse[b].copy_(torch.logsumexp(s, -1)[0])
        o = torch.softmax(s, -1) @ V
        out[rq, :H * D].copy_(o.permute(0, 2, 1, 3).reshape(Sq, H * D).to(out.dtype))
    return out


def attn_bwd(q, k, v, o, dout, lse, delta, dq, dk, dv, B, H, Sq, Skv, D, scale):
    assert o.stride(0) == dout.stride(0) and dq.stride(0) == q.stride(0)
    assert dk.stride(0) == k.stride(0) and dv.stride(0) == v.stride(0)
    for b in range(B):
        rq, rk = slice(b * Sq, (b + 1) * Sq), slice(b * Skv, (b + 1) * Skv)
        Q = _heads(q[rq], 1, Sq, H, D).requires_grad_(True)
        K = _heads(k[rk], 1, Skv, H, D).requires_grad_(True)
        V = _heads(v[rk], 1, Skv, H, D).requires_grad_(True)
        with torch.enable_grad():
            out = torch.softmax((Q @ K.transpose(-1, -2)) * scale, -1) @ V
            (out * _heads(dout[rq], 1, Sq, H, D)).sum().backward()
        for dst, g, S, r in ((dq, Q.grad, Sq, rq), (dk, K.grad, Skv, rk), (dv, V.grad, Skv, rk)):
            dst[r, :H * D].copy_(g.permute(0, 2, 1, 3).reshape(S, H * D).to(dst.dtype))


def geglu_fwd(u, out):
    Fh = u.shape[1] // 2
    uf = u.float()
    out.copy_((uf[:, :Fh] * F.gelu(uf[:, Fh:])).to(out.dtype))     # diffusers GEGLU: hidden * gelu(gate), exact erf
    return out


def geglu_bwd(dgg, u, du):
    Fh = u.shape[1] // 2
    uf = u.float().requires_grad_(True)
    with torch.enable_grad():
        ((uf[:, :Fh] * F.gelu(uf[:, Fh:])) * dgg.float()).sum().backward()
    du.copy_(uf.grad.to(du.dtype))
    return du


def upsample2x_fwd(x, out):
    out.copy_(x.repeat_interleave(2, 1).repeat_interleave(2, 2))
    return out


def upsample2x_bwd(dout, din):
    B, H, W, C = din.shape
    din.copy_(dout.float().view(B, H, 2, W, 2, C).sum((2, 4)).to(din.dtype))
    return din


def conv3x3_c4(x, w, bias, out, sgn=1, round_in=True):
    """4-channel edge convolutions: out[b,h,w,c] = sum_taps x[b, h+sgn*dh, w+sgn*dw, :] . w[c, kh, kw, :]."""
    xf = x.to(BF16).float() if round_in else x.float()
    wt = w.float().permute(0, 3, 1, 2)                 # [C, 4, 3, 3]
    if sgn < 0:
        wt = wt.flip(2, 3)
    y = F.conv2d(xf.permute(0, 3, 1, 2), wt, None if bias is None else bias.float(), padding=1)
    out.copy_(y.permute(0, 2, 3, 1).to(out.dtype))
    return out


def timestep_embed(t, out):
    """diffusers Timesteps(flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = out.shape[1] // 2
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    a = t.float().reshape(-1, 1) * f
    out.copy_(torch.cat([a.cos(), a.sin()], -1).to(out.dtype))
    return out


def cast_f32_bf16(x, out):
    out.copy_(x.to(BF16).view(out.shape))
    return out


def add_bf16(a, b, out):
    out.copy_((a.float() + b.float()).to(out.dtype))
    return out


def install(monkeypatch):
    from pcm_b200 import ops
    g = globals()
    for name in ("groupnorm_fwd", "groupnorm_bwd", "layernorm_fwd", "layernorm_bwd", "attn_fwd", "attn_bwd",
                 "geglu_fwd", "geglu_bwd", "upsample2x_fwd", "upsample2x_bwd", "conv3x3_c4", "timestep_embed",
                 "cast_f32_bf16", "add_bf16"):
        monkeypatch.setattr(ops, name, g[name])
    monkeypatch.setattr(ops, "gemm", interp_gemm)
    monkeypatch.setattr(ops, "wgrad", interp_wgrad)


# ---------------------------------------------------------------------------------------------
# the fused PCM kernels behind ops._call (raw-pointer C ABI, include/pcm_b200.h): restated with the
# pinned oracle functions (oracle/pcm_ref.py == train_pcm_lora_sd15.py:240-341 executed verbatim)
# ---------------------------------------------------------------------------------------------
import ctypes  # noqa: E402


def _raw(ptr, n, ctype, dtype):
    return torch.frombuffer((ctype * n).from_address(ptr), dtype=dtype)


class PcmCalls:
    """Stand-in for ops._call on CPU.  `coef` (the kernels' private [B,16] table) is not modelled: the
    per-step quantities live in this object between pcm_prepare and the consumers."""

    def __init__(self):
        self.s = None

    def __call__(self, name, *args):
        fn = getattr(self, name, None)
        if fn is None:
            raise AssertionError(f"no CPU semantics for {name}")
        return fn(*args)

    def pcm_prepare(self, acp, num_train, num_ddim, inf_idx, multiphase, index, w, B, bf16_mode, coef,
                    start_t, t, end_t):
        from oracle import pcm_ref
        ac = _raw(acp, num_train, ctypes.c_float, torch.float32).clone()
        idx = _raw(index, B, ctypes.c_int64, torch.int64).clone()
        wv = _raw(w, B, ctypes.c_float, torch.float32).clone()
        solver = pcm_ref.DDIMSolverRef(ac.numpy(), num_train, num_ddim)
        st = solver.ddim_timesteps[idx]
        inf = torch.from_numpy(pcm_ref.inference_indices(num_ddim, multiphase)).long()
        assert torch.equal(inf, _raw(inf_idx, multiphase, ctypes.c_int64, torch.int64))
        p = inf[(idx[:, None] >= inf[None, :]).long().sum(1) - 1]
        _raw(start_t, B, ctypes.c_int64, torch.int64).copy_(st)
        _raw(t, B, ctypes.c_int64, torch.int64).copy_(torch.clamp(st - num_train // num_ddim, min=0))
        _raw(end_t, B, ctypes.c_int64, torch.int64).copy_(solver.ddim_timesteps_prev[p])
        self.s = dict(ac=ac, idx=idx, w=wv.to(BF16).float() if bf16_mode else wv, solver=solver, inf=inf,
                      multiphase=multiphase, start_t=st, t=torch.clamp(st - num_train // num_ddim, min=0),
                      alpha=torch.sqrt(ac), sigma=torch.sqrt(1 - ac))

    def pcm_add_noise(self, x, noise, coef, per, B, bf16_mode, out):
        from oracle import pcm_ref
        s = self.s
        xs = _raw(x, B * per, ctypes.c_float, torch.float32).view(B, per)
        ns = _raw(noise, B * per, ctypes.c_float, torch.float32).view(B, per)
        if bf16_mode:
            y = pcm_ref.add_noise(s["ac"], xs.to(BF16), ns.to(BF16), s["start_t"]).float()
        else:
            y = pcm_ref.add_noise(s["ac"], xs, ns, s["start_t"])
        _raw(out, B * per, ctypes.c_float, torch.float32).view(B, per).copy_(y)

    def pcm_teacher_step(self, eps_c, eps_u, noisy, coef, per, B, pred_type, x_prev):
        from oracle import pcm_ref
        s = self.s
        pt = "epsilon" if pred_type == 0 else "v_prediction"
        f = lambda p: _raw(p, B * per, ctypes.c_float, torch.float32).view(B, per)  # noqa: E731
        ec, eu, xn = f(eps_c), f(eps_u), f(noisy)
        x0c = pcm_ref.predicted_origin(ec, s["start_t"], xn, pt, s["alpha"], s["sigma"])
        x0u = pcm_ref.predicted_origin(eu, s["start_t"], xn, pt, s["alpha"], s["sigma"])
        w4 = s["w"].reshape(-1, 1)
        xp = s["solver"].ddim_step(x0c + w4 * (x0c - x0u), ec + w4 * (ec - eu), s["idx"])
        f(x_prev).copy_(xp.float())

    def pcm_lora_refresh(self, master, table, num_entries, total_work, scale, opnd):
        """Table-driven regeneration of the bf16 operand copies (csrc/optim.cu lora_refresh_kernel): row =
        [a_off, b_off, a_fwd, sb_fwd, sb_t, a_t, cin | taps << 32, cout | r << 32, first tile]."""
        tab = _raw(table, num_entries * 9, ctypes.c_int64, torch.int64).view(num_entries, 9).tolist()
        work = 0
        for a_off, b_off, a_fwd, sb_fwd, sb_t, a_t, ci, co, w0 in tab:
            cin, taps, cout, r = ci & 0xffffffff, ci >> 32, co & 0xffffffff, co >> 32
            assert r == 64 and w0 == work
            k = taps * cin
            A = _raw(master + 4 * a_off, r * k, ctypes.c_float, torch.float32).view(r, k).to(BF16)
            sB = (_raw(master + 4 * b_off, cout * r, ctypes.c_float, torch.float32).view(cout, r) * scale).to(BF16)
            o = lambda off, n: _raw(opnd + 2 * off, n, ctypes.c_uint16, torch.int16).view(BF16)  # noqa: E731
            o(a_fwd, r * k).view(r, k).copy_(A)
            o(a_t, r * k).view(cin, taps * r).copy_(A.view(r, taps, cin).permute(2, 1, 0).reshape(cin, taps * r))
            o(sb_fwd, cout * r).view(cout, r).copy_(sB)
            o(sb_t, cout * r).view(r, cout).copy_(sB.t())
            work += k // 64 + cout // 64
        assert work == total_work

    def pcm_grad_sumsq(self, g, n, out):
        _raw(out, 1, ctypes.c_double, torch.float64)[0] = _raw(g, n, ctypes.c_float, torch.float32).double().pow(2).sum()

    def pcm_adamw_clip(self, p, g, m, v, n, state, beta1, beta2, eps, wd, max_norm, inv_world, sumsq, zero_grad):
        """csrc/optim.cu: step counter += 1 on device, clip coefficient of the MEAN gradient from the sum of
        squares of the SUMMED gradient, 1/world folded in, torch.optim.AdamW update, optional zero_grad."""
        f = lambda q: _raw(q, n, ctypes.c_float, torch.float32)  # noqa: E731
        st = _raw(state, 2, ctypes.c_float, torch.float32)
        st[1] += 1.0
        lr, step = st[0].item(), st[1].item()
        norm = float(_raw(sumsq, 1, ctypes.c_double, torch.float64)[0]) ** 0.5 * inv_world
        coef = min(max_norm / (norm + 1e-6), 1.0) if max_norm > 0 else 1.0
        coef *= inv_world
        P, G, M, V = f(p), f(g), f(m), f(v)
        gi = G * coef
        P.mul_(1.0 - lr * wd)
        M.mul_(beta1).add_(gi, alpha=1.0 - beta1)
        V.mul_(beta2).addcmul_(gi, gi, value=1.0 - beta2)
        bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
        P.addcdiv_(M, V.sqrt() / bc2 ** 0.5 + eps, value=-lr / bc1)
        if zero_grad:
            G.zero_()

    def pcm_ema_update(self, targ, src, n, rate):
        t = _raw(targ, n, ctypes.c_float, torch.float32)
        t.mul_(rate).add_(_raw(src, n, ctypes.c_float, torch.float32), alpha=1.0 - rate)

    def pcm_teacher_substep(self, eps_c, eps_u, x_cur, acp, t_cur, t_next, coef, per, B, pred_type, x_next):
        """one DDIM sub-step t_cur -> t_next of the CFG-mixed prediction (t_next < 0: the solver's
        alpha_cumprods[0] entry), float64 like DDIMSolver.ddim_step"""
        from oracle import pcm_ref
        s = self.s
        pt = "epsilon" if pred_type == 0 else "v_prediction"
        f = lambda p: _raw(p, B * per, ctypes.c_float, torch.float32).view(B, per)  # noqa: E731
        ec, eu, xc = f(eps_c), f(eps_u), f(x_cur)
        tc = _raw(t_cur, B, ctypes.c_int64, torch.int64)
        tn = _raw(t_next, B, ctypes.c_int64, torch.int64)
        x0c = pcm_ref.predicted_origin(ec, tc, xc, pt, s["alpha"], s["sigma"])
        x0u = pcm_ref.predicted_origin(eu, tc, xc, pt, s["alpha"], s["sigma"])
        w4 = s["w"].reshape(-1, 1)
        acd = s["ac"].double()
        a_n = torch.where(tn < 0, acd[0], acd[tn.clamp(min=0)]).reshape(-1, 1)
        xn = a_n.sqrt() * (x0c + w4 * (x0c - x0u)) + (1.0 - a_n).sqrt() * (ec + w4 * (ec - eu))
        f(x_next).copy_(xn.float())

    def pcm_loss(self, eps_s, eps_t, noisy, x_prev, coef, per, B, loss_type, huber_c, pred_type, loss_out,
                 d_eps, model_pred, target):
        from oracle import pcm_ref
        s = self.s
        pt = "epsilon" if pred_type == 0 else "v_prediction"
        f = lambda p: _raw(p, B * per, ctypes.c_float, torch.float32).view(B, per)  # noqa: E731
        es = f(eps_s).clone().requires_grad_(True)
        et, xn, xp = f(eps_t), f(noisy), f(x_prev).double()
        c_skip, c_out = [v.reshape(-1, 1) for v in pcm_ref.scalings_for_boundary_conditions_target(s["idx"], s["inf"])]
        with torch.enable_grad():
            x0 = pcm_ref.predicted_origin(es, s["start_t"], xn, pt, s["alpha"], s["sigma"])
            mp, _ = s["solver"].ddim_style_multiphase_pred(x0, es, s["idx"], s["multiphase"])
            x0t = pcm_ref.predicted_origin(et, s["t"], xp, pt, s["alpha"], s["sigma"])
            tg, _ = s["solver"].ddim_style_multiphase_pred(x0t, et, s["idx"], s["multiphase"])
            tg = c_skip * xp + c_out * tg
            d = mp.float() - tg.float()
            loss = (d ** 2).mean() if loss_type == 1 else (torch.sqrt(d ** 2 + huber_c ** 2) - huber_c).mean()
            loss.backward()
        _raw(loss_out, 1, ctypes.c_float, torch.float32).copy_(loss.detach().reshape(1))
        f(d_eps).copy_(es.grad)
        if model_pred:
            f(model_pred).copy_(mp.detach().float())
        if target:
            f(target).copy_(tg.detach().float())


def install_step(monkeypatch):
    from pcm_b200 import ops
    install(monkeypatch)
    monkeypatch.setattr(ops, "_call", PcmCalls())
