"""The kept entry point executed end to end on the GPU (synthetic inputs, narrow UNet so it takes
seconds): `main()` trains, writes checkpoints with optimiser state, rotates them, resumes, and the
written peft adapter reloads; plus the diffusers-shaped model call `unet(...).sample`."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _args(tmp, extra=()):
    from pcm_b200 import config, train_pcm_lora_sd15 as T
    argv = ["--synthetic", "--output_dir", str(tmp), "--train_batch_size", "2", "--resolution", "128",
            "--multiphase", "4", "--loss_type", "huber", "--seed", "5", "--mixed_precision", "bf16",
            "--learning_rate", "1e-3", "--lr_scheduler", "constant_with_warmup", "--lr_warmup_steps", "2",
            "--checkpointing_steps", "2", "--checkpoints_total_limit", "2", "--log_every", "1",
            "--w_min", "4", "--w_max", "5"] + list(extra)
    a = T.parse_args(argv)
    a._cfg = config.TINY
    return T, a


def test_main_trains_checkpoints_and_resumes(cuda, tmp_path):
    from safetensors.torch import load_file
    from pcm_b200 import ops
    ops.deterministic(True, cuda)
    try:
        # straight run: 4 steps
        T, a = _args(tmp_path / "straight", ["--max_train_steps", "4"])
        st_full = T.main(a)
        full = st_full.unet.lora_master.clone()
        # interrupted run: 2 steps (checkpoint-2 written), then resume to 4
        T, a = _args(tmp_path / "resumed", ["--max_train_steps", "2"])
        T.main(a)
        ck = tmp_path / "resumed" / "checkpoint-2"
        assert (ck / "pcm_b200_state.pt").exists() and (ck / "adapter_model.safetensors").exists()
        T, a = _args(tmp_path / "resumed", ["--max_train_steps", "4", "--resume_from_checkpoint", "latest"])
        st_res = T.main(a)
        assert torch.equal(st_res.unet.lora_master, full), "resume is not equivalent to an uninterrupted run"
        assert st_res.opt_state[1].item() == 4.0
        # rotation: limit 2 -> only the newest checkpoints survive
        cks = sorted(d for d in os.listdir(tmp_path / "resumed") if d.startswith("checkpoint"))
        assert cks == ["checkpoint-2", "checkpoint-4"]
        # artefacts reload and match the trained factors
        sd = load_file(str(tmp_path / "resumed" / "adapter_model.safetensors"))
        ref = st_res.unet.lora_state_dict()
        assert len(sd) == len(ref)
        for k, v in ref.items():
            assert torch.equal(sd["base_model.model." + k], v.cpu()), k
        cfgj = json.load(open(tmp_path / "resumed" / "adapter_config.json"))
        assert cfgj["r"] == 64 and cfgj["peft_type"] == "LORA"
        sd2 = load_file(str(tmp_path / "resumed" / "unet_lora" / "pytorch_lora_weights.safetensors"))
        assert all(k.startswith("unet.") for k in sd2)
    finally:
        ops.deterministic(False)


def test_unsupported_flags_fail_loudly():
    from pcm_b200 import train_pcm_lora_sd15 as T
    for bad in (["--use_8bit_adam"], ["--mixed_precision", "no"], ["--gradient_accumulation_steps", "2"],
                ["--lr_scheduler", "nope"]):
        with pytest.raises((ValueError, SystemExit)):
            T.parse_args(["--synthetic"] + bad)


def test_model_call_shape(cuda):
    """unet(sample, timestep, encoder_hidden_states=...).sample with NCHW tensors == the NHWC core."""
    from oracle import unet_ref
    from pcm_b200 import config
    from pcm_b200.modeling import UNet2DConditionModel
    from pcm_b200.unet import UNetB200
    P = unet_ref.init_params(unet_ref.TINY, 0)
    net = UNetB200(config.TINY, P, cuda)
    student, teacher = UNet2DConditionModel(net, use_lora=True), UNet2DConditionModel(net, use_lora=False)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 16, 16, generator=g).to(cuda)
    ctx = torch.randn(2, 77, 64, generator=g).to(cuda)
    ts = torch.tensor([999, 19], device=cuda)
    out = student(x, ts, timestep_cond=None, encoder_hidden_states=ctx.float(), added_cond_kwargs={}).sample
    assert out.shape == x.shape and out.dtype == torch.float32
    ref = unet_ref.UNetRef(unet_ref.TINY, P, use_lora=True, emulate_bf16=True)(x.cpu(), ts.cpu(), ctx.cpu())
    assert ((out.cpu() - ref).norm() / ref.norm()).item() < 3e-2
    t_out = teacher(x, ts, encoder_hidden_states=ctx).sample
    assert (t_out - out).abs().max().item() > 1e-3
    with pytest.raises(RuntimeError):
        student(x.cpu(), ts, encoder_hidden_states=ctx)
    assert student.parameters()[0] is net.lora_master and teacher.parameters() == []


def test_sdxl_entry_point(cuda, tmp_path):
    """train_pcm_lora_sdxl_adv: same flags as the reference script, consistency step only."""
    from pcm_b200 import config, train_pcm_lora_sdxl_adv as TX
    with pytest.raises(ValueError, match="adversarial"):
        TX.parse_args(["--synthetic"])                                   # default --adv_weight 0.1
    a = TX.parse_args(["--synthetic", "--adv_weight", "0", "--output_dir", str(tmp_path), "--train_batch_size", "2",
                       "--resolution", "128", "--multiphase", "4", "--num_ddim_timesteps", "40", "--loss_type", "huber",
                       "--seed", "1", "--mixed_precision", "bf16", "--max_train_steps", "2", "--log_every", "1",
                       "--checkpointing_steps", "100", "--w_min", "4", "--w_max", "5", "--adv_lr", "1e-5"])
    assert a._base_cfg is config.SDXL
    a._cfg = config.TINY_XL                                              # narrow network for the test
    st = TX.main(a)
    assert st.opt_state[1].item() == 2.0 and torch.isfinite(st.loss).all()
    assert (tmp_path / "adapter_model.safetensors").exists()
