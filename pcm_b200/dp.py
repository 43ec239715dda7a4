"""Data-parallel plumbing of the step (one process per GPU, torch.distributed).

The reference wraps the student in DistributedDataParallel through accelerate
(train_pcm_lora_sd15.py:1034): bucketed all-reduce(mean) of the LoRA gradients during backward,
`set_seed(args.seed + process_index)` (:795-797), then clip_grad_norm_ on the reduced gradient.
Here the LoRA gradients already live in ONE flat fp32 buffer, so the exchange is a single
all_reduce(SUM); the 1/world average and the clip coefficient are folded into the AdamW kernel
(`pcm_adamw_clip`), which therefore needs sum-of-squares of the SUMMED gradient.
"""
import math
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def rank_seed(base_seed, rank):
    """accelerate.utils.set_seed(args.seed + accelerator.process_index)."""
    return int(base_seed) + int(rank)


def init_process_group(backend="nccl", device=None):
    """NVLink-only NCCL defaults for one 8xB200 NVSwitch box; rendezvous from MASTER_* env."""
    if backend == "nccl":
        os.environ.setdefault("NCCL_IB_DISABLE", "1")
        os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
        os.environ.setdefault("NCCL_NVLS_ENABLE", "1")
        dist.init_process_group("nccl", device_id=device)
    else:
        dist.init_process_group(backend)
    return dist.group.WORLD


def allreduce_flat_grad(flat_grad, group=None):
    """One collective per step: SUM over ranks of the flat LoRA gradient (269 MB fp32 for SD1.5 r=64)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def folded_coef(sumsq_of_sum, world, max_norm):
    """Scalar applied to the SUMMED gradient so that it equals clip(mean-gradient):
    norm(mean) = sqrt(sumsq)/world ; coef = min(1, max_norm/(norm+1e-6)) / world.
    Mirrors adamw_clip_kernel (csrc/optim.cu); used for logging and by the gloo tests."""
    norm = math.sqrt(float(sumsq_of_sum)) / world
    c = min(max_norm / (norm + 1e-6), 1.0) if max_norm > 0 else 1.0
    return c / world, norm
