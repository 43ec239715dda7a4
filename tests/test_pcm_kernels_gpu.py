"""The fused PCM kernels the training step really executes (`pcm_prepare`, `pcm_teacher_step`,
`pcm_loss`, csrc/pcm_ops.cu), tested in ISOLATION: fixed eps tensors in, results compared with the
golden-pinned restatement of the reference functions (oracle/pcm_ref.py, itself bit-exact against
train_pcm_lora_sd15.py:240-341 executed verbatim, tests/test_oracle.py).

Tolerance 1e-6 relative (fp32 arithmetic, differences are FMA contraction only); the loss gradient
d loss / d eps_student is compared with torch.autograd through the oracle functions.
Covers rows a6 / a7 / a8 / a10 / a11 / a13 / a14 of SURVEY.md section 8, Huber and L2, epsilon and
v_prediction, phase-start rows (c_skip = 1) and interior rows (c_skip = 0), CFG solver on / off.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

HW = 16
PER = 4 * HW * HW
# index values chosen to hit phase starts (0, 12, 25, 37 for 4 phases of 50) and interior steps
INDEX = [0, 5, 12, 13, 24, 25, 37, 49]


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _tables(cuda, multiphase, index, w, bf16_mode=0):
    from pcm_b200 import ops
    from pcm_b200.step import inference_indices, sd15_alphas_cumprod
    B = len(index)
    acp = sd15_alphas_cumprod().to(cuda)
    inf = torch.from_numpy(inference_indices(50, multiphase)).to(cuda)
    idx = torch.tensor(index, device=cuda)
    coef = torch.zeros(B, 16, device=cuda, dtype=torch.float64)
    st, t, et = (torch.zeros(B, device=cuda, dtype=torch.int64) for _ in range(3))
    ops._call("pcm_prepare", acp.data_ptr(), 1000, 50, inf.data_ptr(), multiphase, idx.data_ptr(),
              w.to(cuda).data_ptr(), B, bf16_mode, coef.data_ptr(), st.data_ptr(), t.data_ptr(), et.data_ptr())
    return coef, st, t, et


def _inputs(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda: torch.randn(B, 4, HW, HW, generator=g)
    return dict(eps_s=r(), eps_t=r(), eps_c=r(), eps_u=r(), noisy=r(), w=4.0 + 11.0 * torch.rand(B, generator=g))


def _oracle_teacher(d, index, prediction_type, apply_cfg):
    """T15:1224-1258 through the pinned functions."""
    from oracle import pcm_ref
    ac = pcm_ref.sd15_alphas_cumprod()
    al, sg = torch.sqrt(ac), torch.sqrt(1 - ac)
    solver = pcm_ref.DDIMSolverRef(ac.numpy(), 1000, 50)
    idx = torch.tensor(index)
    start_t = solver.ddim_timesteps[idx]
    x0c = pcm_ref.predicted_origin(d["eps_c"], start_t, d["noisy"], prediction_type, al, sg)
    eu = d["eps_u"] if apply_cfg else d["eps_c"]
    x0u = pcm_ref.predicted_origin(eu, start_t, d["noisy"], prediction_type, al, sg)
    w4 = d["w"].reshape(-1, 1, 1, 1)
    pred_x0 = x0c + w4 * (x0c - x0u)
    pred_noise = d["eps_c"] + w4 * (d["eps_c"] - eu)
    return solver.ddim_step(pred_x0, pred_noise, idx)          # float64


def _oracle_loss(d, x_prev, index, multiphase, prediction_type, loss_type, huber_c, eps_s):
    """T15:1200-1212 + 1269-1293 through the pinned functions; eps_s may require grad."""
    from oracle import pcm_ref
    ac = pcm_ref.sd15_alphas_cumprod()
    al, sg = torch.sqrt(ac), torch.sqrt(1 - ac)
    solver = pcm_ref.DDIMSolverRef(ac.numpy(), 1000, 50)
    idx = torch.tensor(index)
    inf = torch.from_numpy(pcm_ref.inference_indices(50, multiphase)).long()
    start_t = solver.ddim_timesteps[idx]
    t = torch.clamp(start_t - 20, min=0)
    c_skip_s, c_out_s = [pcm_ref.append_dims(x, 4) for x in pcm_ref.scalings_for_boundary_conditions_online(idx, inf)]
    c_skip, c_out = [pcm_ref.append_dims(x, 4) for x in pcm_ref.scalings_for_boundary_conditions_target(idx, inf)]
    x0 = pcm_ref.predicted_origin(eps_s, start_t, d["noisy"], prediction_type, al, sg)
    mp, end_t = solver.ddim_style_multiphase_pred(x0, eps_s, idx, multiphase)
    mp = c_skip_s * d["noisy"] + c_out_s * mp
    x0t = pcm_ref.predicted_origin(d["eps_t"], t, x_prev, prediction_type, al, sg)
    tg, _ = solver.ddim_style_multiphase_pred(x0t, d["eps_t"], idx, multiphase)
    tg = c_skip * x_prev + c_out * tg
    if loss_type == "l2":
        loss = torch.nn.functional.mse_loss(mp.float(), tg.float(), reduction="mean")
    else:
        loss = torch.mean(torch.sqrt((mp.float() - tg.float()) ** 2 + huber_c ** 2) - huber_c)
    return loss, mp, tg, start_t, t, end_t, c_skip.flatten()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("prediction_type", ["epsilon", "v_prediction"])
@pytest.mark.parametrize("apply_cfg", [True, False])
def test_teacher_step_kernel_isolated(cuda, prediction_type, apply_cfg):
    from pcm_b200 import ops
    B = len(INDEX)
    d = _inputs(B, 1)
    coef, st, t, et = _tables(cuda, 4, INDEX, d["w"])
    ref = _oracle_teacher(d, INDEX, prediction_type, apply_cfg)
    ec = _nhwc(d["eps_c"]).to(cuda)
    eu = _nhwc(d["eps_u"]).to(cuda) if apply_cfg else ec
    xn = _nhwc(d["noisy"]).to(cuda)
    out = torch.empty_like(xn)
    ops._call("pcm_teacher_step", ec.data_ptr(), eu.data_ptr(), xn.data_ptr(), coef.data_ptr(), PER, B,
              0 if prediction_type == "epsilon" else 1, out.data_ptr())
    torch.cuda.synchronize()
    assert _rel(_nchw(out).cpu(), ref) <= 1e-6


@pytest.mark.parametrize("prediction_type", ["epsilon", "v_prediction"])
@pytest.mark.parametrize("loss_type", ["huber", "l2"])
@pytest.mark.parametrize("multiphase", [4, 2, 1])
def test_loss_kernel_isolated(cuda, prediction_type, loss_type, multiphase):
    from pcm_b200 import ops
    B = len(INDEX)
    d = _inputs(B, 2)
    huber_c = 1e-3
    coef, st, t, et = _tables(cuda, multiphase, INDEX, d["w"])
    x_prev = _oracle_teacher(d, INDEX, prediction_type, True)      # float64, as the reference feeds it
    eps_s = d["eps_s"].clone().requires_grad_(True)
    loss, mp, tg, start_t, tt, end_t, c_skip = _oracle_loss(d, x_prev, INDEX, multiphase, prediction_type,
                                                            loss_type, huber_c, eps_s)
    loss.backward()
    assert c_skip.sum() > 0 and (1 - c_skip).sum() > 0 or multiphase == 1   # both boundary branches hit
    # integer bookkeeping of pcm_prepare is exact
    assert torch.equal(st.cpu(), start_t) and torch.equal(t.cpu(), tt) and torch.equal(et.cpu(), end_t)
    assert torch.equal(coef[:, 8].cpu().float(), c_skip)
    dev = lambda x: _nhwc(x.float()).to(cuda)
    es, etg, xn, xp = dev(d["eps_s"]), dev(d["eps_t"]), dev(d["noisy"]), dev(x_prev)
    lo = torch.zeros(1, device=cuda)
    de, mpo, tgo = torch.empty_like(es), torch.empty_like(es), torch.empty_like(es)
    ops._call("pcm_loss", es.data_ptr(), etg.data_ptr(), xn.data_ptr(), xp.data_ptr(), coef.data_ptr(), PER, B,
              0 if loss_type == "huber" else 1, huber_c, 0 if prediction_type == "epsilon" else 1,
              lo.data_ptr(), de.data_ptr(), mpo.data_ptr(), tgo.data_ptr())
    torch.cuda.synchronize()
    assert _rel(_nchw(mpo).cpu(), mp.detach()) <= 1e-6
    assert _rel(_nchw(tgo).cpu(), tg.detach()) <= 1e-6
    assert abs(lo.item() - loss.item()) <= 1e-6 * abs(loss.item()), (lo.item(), loss.item())
    # the seed of the UNet backward: d loss / d eps_student vs autograd (fp32 graph: 1e-4)
    assert _rel(_nchw(de).cpu(), eps_s.grad) <= 1e-4


def test_prepare_bf16_mode_matches_reference_casts(cuda):
    """bf16_mode: w is rounded like `w.to(latents.dtype)` (T15:1185) and the add_noise coefficients
    follow add_noise's casts (S15:510-523)."""
    from oracle import pcm_ref
    d = _inputs(len(INDEX), 3)
    coef, st, t, et = _tables(cuda, 4, INDEX, d["w"], bf16_mode=1)
    assert torch.equal(coef[:, 9].cpu().float(), d["w"].bfloat16().float())
    ac = pcm_ref.sd15_alphas_cumprod().to(torch.bfloat16)
    sa = (ac[st.cpu()] ** 0.5).float()
    so = ((1 - ac[st.cpu()]) ** 0.5).float()
    assert torch.equal(coef[:, 10].cpu().float(), sa) and torch.equal(coef[:, 11].cpu().float(), so)


def test_teacher_substep_single_equals_teacher_step(cuda):
    """pcm_teacher_substep with ONE sub-step over the whole interval == pcm_teacher_step, bit for bit
    (the reference behaviour is the k = 1 case of the opt-in multi-substep solve)."""
    from pcm_b200 import ops
    from pcm_b200.step import sd15_alphas_cumprod
    B = len(INDEX)
    d = _inputs(B, 5)
    coef, st, t, et = _tables(cuda, 4, INDEX, d["w"])
    ec, eu, xn = (_nhwc(d[k]).to(cuda) for k in ("eps_c", "eps_u", "noisy"))
    a, b = torch.empty_like(xn), torch.empty_like(xn)
    ops._call("pcm_teacher_step", ec.data_ptr(), eu.data_ptr(), xn.data_ptr(), coef.data_ptr(), PER, B, 0, a.data_ptr())
    acp = sd15_alphas_cumprod().to(cuda)
    t_next = st - 20
    ops._call("pcm_teacher_substep", ec.data_ptr(), eu.data_ptr(), xn.data_ptr(), acp.data_ptr(), st.data_ptr(),
              t_next.data_ptr(), coef.data_ptr(), PER, B, 0, b.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(a, b)
