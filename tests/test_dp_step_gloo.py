"""World-size-2 `gloo` run of the PRODUCT data-parallel step on CPU (kernels replaced by their torch
semantics, tests/ops_interp.py): PCMTrainStep with a process group - bucketed all-reduce(SUM) launched
from inside the backward walk (dp.GradReducer), 1/world and the clip folded into the AdamW launch.  Two
ranks with DIFFERENT batches (seed + rank, T15:795-797) must end with identical LoRA parameters, equal to
one process applying clip_grad_norm_ + AdamW to the MEAN of the two local gradients - what DDP + accelerate
do in the reference (train_pcm_lora_sd15.py:1034, 1296-1301).

CPU twin of tests/test_dp_nccl_gpu.py (which needs 2 GPUs and is skipped on a single-GPU box)."""
import os
import socket
import sys

import pytest
import torch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _worker(rank, world, port, out, overlap):
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      PCM_DP_OVERLAP=overlap)
    torch.set_num_threads(2)
    import torch.distributed as dist
    import ops_interp
    from gemm_interp import refresh_operands
    from oracle import pcm_ref, unet_ref
    from pcm_b200 import config, dp, ops
    from pcm_b200.step import PCMTrainStep

    class MP:                                  # minimal monkeypatch stand-in inside the spawned process
        @staticmethod
        def setattr(obj, name, val):
            setattr(obj, name, val)

    pg = dp.init_process_group("gloo")
    B, hw, mp_ = 2, 8, 4
    P = unet_ref.init_params(unet_ref.TINY, 0)                                  # replicated weights
    batch = pcm_ref.make_batch(unet_ref.TINY, B, hw, seed=dp.rank_seed(11, rank))     # per-rank data
    args = (_nhwc(batch["latents"]), _nhwc(batch["noise"]), batch["index"], batch["w"],
            batch["prompt_embeds"].to(torch.bfloat16), batch["uncond_prompt_embeds"].to(torch.bfloat16))
    kw = dict(batch=B, height=hw, width=hw, multiphase=mp_, lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0)

    def build(**extra):
        ops.DRY_RUN = []
        try:
            st = PCMTrainStep(config.TINY, P, "cpu", **kw, **extra)
        finally:
            ops.DRY_RUN = None
        refresh_operands(st.unet)
        return st

    solo, st = build(), build(process_group=pg)
    ops_interp.install_step(MP)
    solo.load_inputs(*args)
    solo.forward_backward()                                    # this rank's local gradient, no collective
    g_local = solo.unet.lora_grad.clone()
    p0 = solo.unet.lora_master.clone()
    assert len(st.reducer.buckets) > 1 and st.world == world
    st.load_inputs(*args)
    st.step()                                                  # the product data-parallel iteration
    gathered = [torch.zeros_like(g_local) for _ in range(world)]
    dist.all_gather(gathered, g_local)
    master = st.unet.lora_master.clone()
    masters = [torch.zeros_like(master) for _ in range(world)]
    dist.all_gather(masters, master)
    if rank == 0:
        out["masters_equal"] = all(torch.equal(masters[0], m) for m in masters[1:])
        out["grads_differ"] = not torch.equal(gathered[0], gathered[1])
        mean = (sum(g.double() for g in gathered) / world).float()
        params = {"w": p0.clone()}
        pcm_ref.clip_and_adamw_ref(params, {"w": mean}, {}, lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0)
        out["max_abs_err"] = (master - params["w"]).abs().max().item()
        out["update_size"] = (params["w"] - p0).abs().max().item()
        out["grad_zeroed"] = st.unet.lora_grad.abs().max().item() == 0.0
        out["step_count"] = st.opt_state[1].item()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", ["1", "0"])
def test_two_rank_product_step_matches_mean_gradient_adamw(overlap):
    """overlap=1: bucketed all-reduces launched from the backward walk; 0: one flat all-reduce after it."""
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out, overlap), nprocs=world, join=True)
    assert out["grads_differ"], "the two ranks saw the same batch"
    assert out["masters_equal"], "ranks ended with different parameters"
    assert out["update_size"] > 1e-5 and out["grad_zeroed"] and out["step_count"] == 1.0
    assert out["max_abs_err"] <= 1e-6 + 1e-3 * out["update_size"], dict(out)
