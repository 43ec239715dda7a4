"""Flow-matching solver / scheduler mirrors (pcm_b200/fm_schedulers.py) against golden vectors made
by executing the reference classes verbatim (tests/golden/make_fm_golden.py -> fm_math.pt):
host-side tables, state and error behaviour on the CPU; the step arithmetic on the GPU (bit-exact for
the fp32 scheduler steps, 1e-12 for the float64 EulerSolver results)."""
import os

import pytest
import torch

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fm_math.pt"))


def _cls(name):
    from pcm_b200 import fm_schedulers as F
    return F.PCMFMDeterministicScheduler if name == "det" else F.PCMFMStochasticScheduler


@pytest.mark.parametrize("name", ["det", "sto"])
@pytest.mark.parametrize("shift", [1.0, 3.0])
@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_scheduler_tables_and_state(name, shift, n):
    g = GOLD[f"{name}_shift{shift}_n{n}"]
    s = _cls(name)(num_train_timesteps=1000, shift=shift, pcm_timesteps=50)
    assert torch.equal(s.sigmas, g["sigmas"]) and torch.equal(s.timesteps, g["timesteps0"])
    assert s.sigma_min == g["sigma_min"] and s.sigma_max == g["sigma_max"]
    assert s.config.num_train_timesteps == 1000 and s.config.shift == shift and len(s) == 1000
    s.set_timesteps(n)
    assert torch.equal(s.timesteps, g["timesteps"]) and torch.equal(s.sigmas_, g["sigmas_"])
    assert s.step_index is None and s.begin_index is None
    assert s.index_for_timestep(s.timesteps[0]) == 0
    with pytest.raises(ValueError):
        s.step(torch.zeros(1), 3, torch.zeros(1))                       # integer timesteps are rejected
    with pytest.raises(ValueError):
        s.step(torch.zeros(1), torch.tensor(3), torch.zeros(1))
    with pytest.raises(RuntimeError):
        s.step(torch.zeros(1, 4), s.timesteps[0], torch.zeros(1, 4))    # CPU tensors: no CPU fallback


def test_euler_solver_tables():
    from pcm_b200.fm_schedulers import EulerSolver
    e = GOLD["euler"]
    sv = EulerSolver(GOLD["sigmas_train"].numpy(), 1000, 50)
    assert torch.equal(sv.euler_timesteps, e["euler_timesteps"])
    assert torch.equal(sv.euler_timesteps_prev, e["euler_timesteps_prev"])
    assert torch.equal(sv.sigmas, e["sigmas"]) and sv.sigmas.dtype == e["sigmas"].dtype
    assert torch.equal(sv.sigmas_prev, e["sigmas_prev"]) and sv.sigmas_prev.dtype == torch.float64
    assert sv.step_ratio == 20


@pytest.mark.gpu
def test_euler_solver_steps_gpu(cuda):
    from pcm_b200.fm_schedulers import EulerSolver
    sv = EulerSolver(GOLD["sigmas_train"].numpy(), 1000, 50).to(cuda)
    x, v, idx = GOLD["x"].to(cuda), GOLD["v"].to(cuda), GOLD["idx"].to(cuda)
    out = sv.euler_step(x, v, idx)
    assert out.dtype == torch.float64
    assert torch.allclose(out.cpu(), GOLD["euler_step"], rtol=1e-12, atol=1e-12)
    for mp in (1, 2, 4):
        for tgt in (False, True):
            xp, end = sv.euler_style_multiphase_pred(x, v, idx, mp, is_target=tgt)
            gx, gend = GOLD[f"euler_mp{mp}_{int(tgt)}"]
            assert torch.equal(end.cpu(), gend)
            assert torch.allclose(xp.cpu(), gx, rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["det", "sto"])
@pytest.mark.parametrize("shift", [1.0, 3.0])
@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_scheduler_steps_gpu(cuda, name, shift, n):
    g = GOLD[f"{name}_shift{shift}_n{n}"]
    s = _cls(name)(num_train_timesteps=1000, shift=shift, pcm_timesteps=50)
    s.set_timesteps(n, device=cuda)
    x, v = GOLD["x"].to(cuda), GOLD["v"].to(cuda)
    cur = x.clone()
    for i, ts in enumerate(s.timesteps):
        ref, z = g["steps"][i]
        cur = s.step(v * (1 + 0.1 * i), ts, cur, noise=z.to(cuda)).prev_sample
        assert cur.dtype == torch.float32
        # same fp32 operation order as the reference, no fma contraction: bit-exact
        assert torch.equal(cur.cpu(), ref), (name, shift, n, i, (cur.cpu() - ref).abs().max().item())
    assert s.step_index == n
    s2 = _cls(name)(num_train_timesteps=1000, shift=shift, pcm_timesteps=50)
    s2.set_timesteps(n, device=cuda)
    sn = s2.scale_noise(x, s2.timesteps[0], GOLD["noise"].to(cuda))
    assert torch.equal(sn.cpu(), g["scale_noise"])
