"""CPU checks of host-side logic that the GPU path relies on:
  * the oracle's clip + AdamW restatement pinned against torch.nn.utils.clip_grad_norm_ +
    torch.optim.AdamW (what train_pcm_lora_sd15.py:972-991, 1297-1301 call);
  * product layer table / synthetic weights == the oracle's (they are duplicated on purpose so the
    product never imports the oracle);
  * learning-rate schedules against torch LambdaLR driven by the same lambdas;
  * gradient bucket boundaries of the overlapped all-reduce;
  * checkpoint rotation / resume discovery of the training entry point."""
import math
import os

import pytest
import torch


def test_clip_and_adamw_ref_matches_torch_optim():
    from oracle import pcm_ref
    g = torch.Generator().manual_seed(0)
    shapes = {"a.lora_A.weight": (64, 320), "a.lora_B.weight": (320, 64), "b.lora_A.weight": (64, 3, 3, 8)}
    params = {k: torch.randn(s, generator=g) for k, s in shapes.items()}
    tp = {k: torch.nn.Parameter(v.clone()) for k, v in params.items()}
    opt = torch.optim.AdamW(list(tp.values()), lr=1e-2, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    state = {}
    for it in range(4):
        scale = [3.0, 0.01, 1.0, 0.3][it]       # norms above and below max_grad_norm
        grads = {k: torch.randn(s, generator=g) * scale for k, s in shapes.items()}
        for k in tp:
            tp[k].grad = grads[k].clone()
        tn = torch.nn.utils.clip_grad_norm_(list(tp.values()), 1.0)
        opt.step()
        total = pcm_ref.clip_and_adamw_ref(params, grads, state, lr=1e-2, weight_decay=1e-2, max_grad_norm=1.0)
        # fp32 reductions in a different order (torch's foreach norm is threaded): a few ulp, not 1e-6 exactly
        assert abs(total.item() - tn.item()) <= 1e-5 * tn.item()
        for k in params:
            assert torch.allclose(params[k], tp[k].detach(), rtol=1e-5, atol=1e-6), (it, k)


def test_product_tables_equal_oracle_tables():
    from oracle import unet_ref
    from pcm_b200 import config, weights
    for name in ("SD15", "TINY", "SDXL", "TINY_XL"):
        o, p = getattr(unet_ref, name), getattr(config, name)
        assert unet_ref.layer_table(o) == config.layer_table(p)
    for name in ("TINY", "TINY_XL"):
        a = unet_ref.init_params(getattr(unet_ref, name), 3)
        b = weights.synthetic_state_dict(getattr(config, name), 3)
        assert a.keys() == b.keys()
        assert all(torch.equal(a[k], b[k]) for k in a)
    # SDXL inventory: 2.567 B base parameters, transformer depth 1 / 2 / 10
    tab = config.layer_table(config.SDXL)
    assert sum(1 for t in tab if ".transformer_blocks.9.attn1.to_q" in t[0]) == 6   # depth-10 stacks: 2 down + mid + 3 up
    assert any(t[0] == "add_embedding.linear_1" and t[2] == 2816 for t in tab)


@pytest.mark.parametrize("name", ["constant", "constant_with_warmup", "linear", "cosine", "cosine_with_restarts",
                                  "polynomial", "piecewise_constant"])
def test_lr_schedules_match_lambdalr(name):
    from pcm_b200 import lr_schedules as S
    base, warm, total = 1e-3, 5, 40
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=base)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: S.lr_multiplier(name, s, warm, total, base))
    for k in range(50):
        assert math.isclose(opt.param_groups[0]["lr"], S.lr_at(name, base, k, warm, total), rel_tol=1e-12, abs_tol=1e-18)
        opt.step()
        sched.step()
    # accelerate steps the scheduler once per process per optimiser step
    assert S.lr_at("constant_with_warmup", base, 1, 8, total, num_processes=4) == base * 4 / 8
    assert S.lr_at("constant_with_warmup", base, 0, 8, total) == 0.0       # LambdaLR starts at lambda(0)
    with pytest.raises(ValueError):
        S.lr_multiplier("nope", 0)


def test_bucket_bounds():
    from pcm_b200 import dp
    offs = [0, 100, 250, 400, 700, 900]
    b = dp.bucket_bounds(offs, 1000, 4)
    assert b[0][0] == 0 and b[-1][1] == 1000
    assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
    assert all(lo in offs for lo, _ in b)
    assert dp.bucket_bounds(offs, 1000, 1) == [(0, 1000)]
    assert dp.bucket_bounds([], 10, 4) == [(0, 10)]

    class FakeReducer(dp.GradReducer):
        def __init__(self):
            self.world, self.buckets, self.launched = 2, b, []
            self._next, self._work = len(b) - 1, []

        def ready(self, lo):
            while self._next >= 0 and self.buckets[self._next][0] >= lo:
                self.launched.append(self.buckets[self._next])
                self._next -= 1
    r = FakeReducer()
    r.ready(950)
    assert r.launched == []
    r.ready(700)
    assert r.launched == [bb for bb in reversed(b) if bb[0] >= 700]
    r.ready(0)
    assert sorted(r.launched) == b


def test_checkpoint_rotation_and_resume_discovery(tmp_path):
    from pcm_b200 import train_pcm_lora_sd15 as T
    out = str(tmp_path)
    for s in (10, 20, 30):
        os.makedirs(os.path.join(out, f"checkpoint-{s}"))
    T.rotate_checkpoints(out, 3)            # keeps at most limit - 1 before the new save
    assert sorted(os.listdir(out)) == ["checkpoint-20", "checkpoint-30"]
    assert T.find_resume_path(out, "latest") == "checkpoint-30"
    assert T.find_resume_path(out, "/some/where/checkpoint-20") == "checkpoint-20"
    assert T.find_resume_path(str(tmp_path / "empty"), "latest") is None


def test_parameter_counts_match_published_checkpoints():
    """External anchor for the (otherwise unpinned) UNet restatement: the layer inventory reproduces the
    parameter counts of the published checkpoints exactly - runwayml/stable-diffusion-v1-5 UNet
    859,520,964 and stabilityai/stable-diffusion-xl-base-1.0 UNet 2,567,463,684 - and the peft wrapping
    rule gives 278 LoRA modules / 67,252,224 LoRA parameters at r = 64 for SD1.5."""
    from pcm_b200 import config

    def count(cfg):
        n = lora = mods = 0
        for name, kind, cin, cout, k in config.layer_table(cfg):
            if kind in ("gn", "ln"):
                n += 2 * cout
                continue
            taps = k * k if kind == "conv" else 1
            n += cout * cin * taps + (0 if kind == "linear_nobias" else cout)
            if config.is_lora_target(name):
                lora += cfg.lora_rank * (cin * taps + cout)
                mods += 1
        return n, lora, mods
    n15, l15, m15 = count(config.SD15)
    assert n15 == 859_520_964 and l15 == 67_252_224 and m15 == 278
    nxl, lxl, mxl = count(config.SDXL)
    assert nxl == 2_567_463_684 and mxl == 788
