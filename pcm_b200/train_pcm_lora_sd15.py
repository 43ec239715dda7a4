#!/usr/bin/env python
"""`train_pcm_lora_sd15` entry point on the B200 path.

Keeps the reference script's command line (flag names / defaults of
/root/reference/code/text_to_image_sd15/train_pcm_lora_sd15.py:381-735) so existing launch recipes
(`train_pcm_lora_sd15.sh`) keep working; the per-iteration hot path runs through libpcm_b200.so
(PCMTrainStep).  What is NOT here (out of scope): the image dataset, VAE and CLIP
encoders, validation image logging, hub upload.  Inputs to the step are therefore either
  --synthetic            seeded random latents / text embeddings (benchmark + parity mode), or
  --latent_cache DIR     *.pt files holding {"latents" [B,4,h,w], "prompt_embeds" [B,77,768]}
                         produced upstream by the reference's own VAE/CLIP preprocessing.
Launch: `python -m pcm_b200.train_pcm_lora_sd15 ...` or
`python -m torch.distributed.run --nproc-per-node N -m pcm_b200.train_pcm_lora_sd15 ...`.
"""
import argparse
import glob
import json
import os
import time

import torch

from . import config, weights
from .step import PCMTrainStep


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="PCM-LoRA distillation (SD1.5) on B200")
    # ---- flags of the reference script, verbatim names / defaults ----
    p.add_argument("--pretrained_teacher_model", type=str, default=None,
                   help="directory with unet/diffusion_pytorch_model.safetensors; omitted -> seeded random init")
    p.add_argument("--pretrained_vae_model_name_or_path", type=str, default=None)
    p.add_argument("--teacher_revision", type=str, default=None)
    p.add_argument("--revision", type=str, default=None)
    p.add_argument("--output_dir", type=str, default="lcm-xl-distilled")
    p.add_argument("--cache_dir", type=str, default=None)
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--logging_dir", type=str, default="logs")
    p.add_argument("--report_to", type=str, default="tensorboard")
    p.add_argument("--checkpointing_steps", type=int, default=500)
    p.add_argument("--checkpoints_total_limit", type=int, default=None)
    p.add_argument("--resume_from_checkpoint", type=str, default=None)
    p.add_argument("--resolution", type=int, default=512)
    p.add_argument("--center_crop", default=False, action="store_true")
    p.add_argument("--random_flip", action="store_true")
    p.add_argument("--dataloader_num_workers", type=int, default=8)
    p.add_argument("--train_batch_size", type=int, default=16)
    p.add_argument("--num_train_epochs", type=int, default=100)
    p.add_argument("--max_train_steps", type=int, default=None)
    p.add_argument("--max_train_samples", type=int, default=None)
    p.add_argument("--learning_rate", type=float, default=1e-4)
    p.add_argument("--scale_lr", action="store_true", default=False)
    p.add_argument("--lr_scheduler", type=str, default="constant")
    p.add_argument("--lr_warmup_steps", type=int, default=500)
    p.add_argument("--gradient_accumulation_steps", type=int, default=1)
    p.add_argument("--use_8bit_adam", action="store_true")
    p.add_argument("--adam_beta1", type=float, default=0.9)
    p.add_argument("--adam_beta2", type=float, default=0.999)
    p.add_argument("--adam_weight_decay", type=float, default=1e-2)
    p.add_argument("--adam_epsilon", type=float, default=1e-08)
    p.add_argument("--max_grad_norm", default=1.0, type=float)
    p.add_argument("--proportion_empty_prompts", type=float, default=0)
    p.add_argument("--w_min", type=float, default=5.0)
    p.add_argument("--w_max", type=float, default=15.0)
    p.add_argument("--num_ddim_timesteps", type=int, default=50)
    p.add_argument("--loss_type", type=str, default="l2", choices=["l2", "huber"])
    p.add_argument("--huber_c", type=float, default=0.001)
    p.add_argument("--lora_rank", type=int, default=64)
    p.add_argument("--mixed_precision", type=str, default=None, choices=["no", "fp16", "bf16"])
    p.add_argument("--allow_tf32", action="store_true")
    p.add_argument("--cast_teacher_unet", action="store_true")
    p.add_argument("--enable_xformers_memory_efficient_attention", action="store_true")
    p.add_argument("--gradient_checkpointing", action="store_true")
    p.add_argument("--local_rank", type=int, default=-1)
    p.add_argument("--validation_steps", type=int, default=200)
    p.add_argument("--push_to_hub", action="store_true")
    p.add_argument("--hub_token", type=str, default=None)
    p.add_argument("--hub_model_id", type=str, default=None)
    p.add_argument("--tracker_project_name", type=str, default="text2image-fine-tune")
    p.add_argument("--not_apply_cfg_solver", action="store_true")
    p.add_argument("--multiphase", default=8, type=int)
    # ---- additions of this implementation ----
    p.add_argument("--synthetic", action="store_true", help="seeded synthetic latents / text embeddings")
    p.add_argument("--latent_cache", type=str, default=None)
    p.add_argument("--no_cuda_graph", action="store_true")
    p.add_argument("--log_every", type=int, default=10)
    args = p.parse_args(argv)
    env_local_rank = int(os.environ.get("LOCAL_RANK", -1))   # same override as the reference
    if env_local_rank != -1 and env_local_rank != args.local_rank:
        args.local_rank = env_local_rank
    if args.proportion_empty_prompts < 0 or args.proportion_empty_prompts > 1:
        raise ValueError("`--proportion_empty_prompts` must be in the range [0, 1].")
    if args.gradient_accumulation_steps != 1:
        raise ValueError("pcm_b200 runs one optimiser step per iteration (all reference recipes use 1)")
    if args.use_8bit_adam:
        raise ValueError("--use_8bit_adam (bitsandbytes) is not provided: 180 GB HBM holds fp32 AdamW state")
    if args.mixed_precision == "fp16":
        raise ValueError("pcm_b200 computes in bf16 (tcgen05 kind::f16 with bf16 operands); use --mixed_precision bf16")
    return args


def _lr_at(args, step):
    if args.lr_scheduler == "constant":
        return args.learning_rate
    if args.lr_scheduler == "constant_with_warmup":
        return args.learning_rate * min(1.0, (step + 1) / max(1, args.lr_warmup_steps))
    raise ValueError(f"lr scheduler {args.lr_scheduler} not supported")


def main(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = max(args.local_rank, 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
        pg = torch.distributed.group.WORLD
    if args.seed is not None:
        torch.manual_seed(args.seed + rank)        # set_seed(args.seed + process_index)
    cfg = config.UNetConfig(lora_rank=args.lora_rank)
    if args.pretrained_teacher_model:
        from safetensors.torch import load_file
        sd = load_file(os.path.join(args.pretrained_teacher_model, "unet", "diffusion_pytorch_model.safetensors"))
        sd.update({k: v for k, v in weights.synthetic_state_dict(cfg, args.seed or 0, lora_b_std=0.0).items()
                   if ".lora_" in k})              # peft init: A kaiming-uniform, B zeros
    else:
        sd = weights.synthetic_state_dict(cfg, args.seed or 0)
    hw = args.resolution // 8
    B = args.train_batch_size
    lr = args.learning_rate * (B * world if args.scale_lr else 1)
    args.learning_rate = lr
    st = PCMTrainStep(cfg, sd, dev, batch=B, height=hw, width=hw, multiphase=args.multiphase,
                      num_ddim_timesteps=args.num_ddim_timesteps, loss_type=args.loss_type,
                      huber_c=args.huber_c, lr=lr, betas=(args.adam_beta1, args.adam_beta2),
                      adam_eps=args.adam_epsilon, weight_decay=args.adam_weight_decay,
                      max_grad_norm=args.max_grad_norm, apply_cfg_solver=not args.not_apply_cfg_solver,
                      process_group=pg)
    del sd
    files = sorted(glob.glob(os.path.join(args.latent_cache, "*.pt"))) if args.latent_cache else []
    if not files and not args.synthetic:
        raise SystemExit("need --synthetic or --latent_cache (the image/VAE/CLIP pipeline is out of scope)")
    gen = torch.Generator().manual_seed((args.seed or 0) + rank)
    uncond = torch.zeros(B, 77, cfg.cross_attention_dim)

    def next_batch(i):
        if files:
            d = torch.load(files[(i * world + rank) % len(files)])
            lat, pe = d["latents"].float(), d["prompt_embeds"].float()
            unc = d.get("uncond_prompt_embeds", uncond)
        else:
            lat = torch.randn(B, 4, hw, hw, generator=gen)
            pe = torch.randn(B, 77, cfg.cross_attention_dim, generator=gen)
            unc = uncond
        noise = torch.randn(B, 4, hw, hw, generator=gen)
        index = torch.randint(0, args.num_ddim_timesteps, (B,), generator=gen)
        w = (args.w_max - args.w_min) * torch.rand(B, generator=gen) + args.w_min
        nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous()
        return nhwc(lat), nhwc(noise), index, w, pe.bfloat16(), unc.bfloat16()

    st.load_inputs(*next_batch(0))
    if not args.no_cuda_graph and world == 1:
        st.capture()
    max_steps = args.max_train_steps or 1000
    os.makedirs(args.output_dir, exist_ok=True)
    t0 = time.time()
    for step in range(max_steps):
        st.load_inputs(*next_batch(step))
        st.set_lr(_lr_at(args, step))
        st.step()
        if rank == 0 and (step + 1) % args.log_every == 0:
            print(json.dumps({"step": step + 1, "loss": st.loss.item(), "lr": _lr_at(args, step),
                              "steps_per_s": (step + 1) / (time.time() - t0)}), flush=True)
        if rank == 0 and (step + 1) % args.checkpointing_steps == 0:
            save_lora(st, cfg, os.path.join(args.output_dir, f"checkpoint-{step + 1}"))
    if rank == 0:
        save_lora(st, cfg, args.output_dir)
    if world > 1:
        torch.distributed.destroy_process_group()


def save_lora(st, cfg, out_dir):
    """peft adapter + diffusers `unet_lora/pytorch_lora_weights.safetensors` (T15:924-928, 1378-1382)."""
    from safetensors.torch import save_file
    os.makedirs(os.path.join(out_dir, "unet_lora"), exist_ok=True)
    lora = {k: v.cpu().contiguous() for k, v in st.unet.lora_state_dict().items()}
    save_file(weights.to_peft_keys(lora), os.path.join(out_dir, "adapter_model.safetensors"))
    save_file({"unet." + k: v for k, v in lora.items()}, os.path.join(out_dir, "unet_lora", "pytorch_lora_weights.safetensors"))
    json.dump({"r": cfg.lora_rank, "lora_alpha": cfg.lora_alpha, "target_modules": list(config.LORA_TARGETS),
               "peft_type": "LORA"}, open(os.path.join(out_dir, "adapter_config.json"), "w"))


if __name__ == "__main__":
    main(parse_args())
