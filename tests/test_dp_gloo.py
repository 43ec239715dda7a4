"""world_size-2 gloo test (CPU) of the data-parallel host logic: one flat all_reduce(SUM) + the folded
1/world / clip coefficient gives every rank the bit-identical update that DDP's all-reduce(mean) +
clip_grad_norm_ + AdamW (train_pcm_lora_sd15.py:1034, 1296-1301) would give."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from pcm_b200 import dp
    from oracle import pcm_ref
    dp.init_process_group("gloo")
    n = 4096
    g = torch.Generator().manual_seed(dp.rank_seed(1234, rank))
    grad = torch.randn(n, generator=g) * 0.05          # per-rank gradient (different data shard)
    p0 = torch.randn(n, generator=torch.Generator().manual_seed(7))   # replicated parameters
    flat = grad.clone()
    dp.allreduce_flat_grad(flat)
    coef, norm = dp.folded_coef((flat.double() ** 2).sum().item(), world, 1.0)
    # AdamW on coef * summed gradient (what pcm_adamw_clip does)
    params = {"w": p0.clone()}
    pcm_ref.clip_and_adamw_ref(params, {"w": flat * coef}, {}, lr=1e-2, weight_decay=1e-2, max_grad_norm=0.0)
    # reference semantics: mean gradient, clip_grad_norm_, AdamW
    gathered = [torch.zeros(n) for _ in range(world)]
    dist.all_gather(gathered, grad)
    mean = sum(gathered) / world
    ref = {"w": p0.clone()}
    total = pcm_ref.clip_and_adamw_ref(ref, {"w": mean}, {}, lr=1e-2, weight_decay=1e-2, max_grad_norm=1.0)
    out[rank] = (params["w"], ref["w"], norm, float(total))
    dist.destroy_process_group()


def test_flat_allreduce_matches_ddp_mean_clip():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    (p0, r0, n0, t0), (p1, r1, n1, t1) = out[0], out[1]
    assert torch.equal(p0, p1)                       # every rank applies the identical update
    assert torch.allclose(p0, r0, rtol=1e-5, atol=1e-7)
    assert abs(n0 - t0) < 1e-5 * max(1.0, t0)


def test_rank_seeds_differ():
    from pcm_b200 import dp
    assert dp.rank_seed(5, 0) != dp.rank_seed(5, 1)
    assert dp.folded_coef(4.0, 2, 0.0) == (0.5, 1.0)


def _worker_reducer(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from pcm_b200 import dp
    dp.init_process_group("gloo")
    n = 1000
    grad = torch.randn(n, generator=torch.Generator().manual_seed(dp.rank_seed(77, rank)))
    local = grad.clone()
    red = dp.GradReducer(grad, [0, 100, 250, 400, 700, 900], num_buckets=4)
    assert red.world == world and len(red.buckets) == 4
    red.start()
    # the backward pass reports completed offsets from the END of the buffer, block by block
    for lo in (900, 700, 400, 250, 100, 0):
        red.ready(lo)
    red.finish()
    gathered = [torch.zeros(n) for _ in range(world)]
    dist.all_gather(gathered, local)
    out[rank] = (grad.clone(), sum(gathered))
    dist.destroy_process_group()


def test_grad_reducer_buckets_sum_to_flat_allreduce():
    """The product's bucketed, backward-ordered all-reduce (dp.GradReducer, used by PCMTrainStep in eager
    mode) gives exactly the flat SUM on every rank."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_reducer, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        got, ref = out[r]
        assert torch.equal(got, ref)
    assert torch.equal(out[0][0], out[1][0])
