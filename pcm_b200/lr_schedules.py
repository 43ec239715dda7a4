"""Learning-rate schedules of `diffusers.optimization.get_scheduler` (called at
train_pcm_lora_sd15.py:1026-1031 with name = --lr_scheduler, num_warmup_steps, num_training_steps),
as pure functions step -> multiplier of the base learning rate.

diffusers builds a torch LambdaLR: the multiplier used by optimiser step k (0-based) is
lr_lambda(k') where k' is the number of scheduler.step() calls so far.  accelerate's
AcceleratedScheduler (accelerate==0.27.2, the pinned version) calls scheduler.step() once per
PROCESS for every optimiser step when split_batches=False (the script's setting), so under N-GPU data
parallelism the reference's schedule advances N ticks per iteration: k' = k * num_processes.
`lr_at` reproduces that.
"""
import math

SCHEDULES = ("linear", "cosine", "cosine_with_restarts", "polynomial", "constant", "constant_with_warmup",
             "piecewise_constant")


def _constant(step, warmup, total):
    return 1.0


def _constant_with_warmup(step, warmup, total):
    if step < warmup:
        return float(step) / float(max(1.0, warmup))
    return 1.0


def _linear(step, warmup, total):
    if step < warmup:
        return float(step) / float(max(1, warmup))
    return max(0.0, float(total - step) / float(max(1, total - warmup)))


def _cosine(step, warmup, total, num_cycles=0.5):
    if step < warmup:
        return float(step) / float(max(1, warmup))
    progress = float(step - warmup) / float(max(1, total - warmup))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))


def _cosine_with_restarts(step, warmup, total, num_cycles=1):
    if step < warmup:
        return float(step) / float(max(1, warmup))
    progress = float(step - warmup) / float(max(1, total - warmup))
    if progress >= 1.0:
        return 0.0
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * ((float(num_cycles) * progress) % 1.0))))


def _polynomial(step, warmup, total, lr_init, lr_end=1e-7, power=1.0):
    if not (lr_init > lr_end):
        raise ValueError(f"lr_end ({lr_end}) must be be smaller than initial lr ({lr_init})")
    if step < warmup:
        return float(step) / float(max(1, warmup))
    if step > total:
        return lr_end / lr_init
    lr_range = lr_init - lr_end
    decay_steps = total - warmup
    pct_remaining = 1 - (step - warmup) / decay_steps
    return (lr_range * pct_remaining ** power + lr_end) / lr_init


def _piecewise_constant(step, step_rules="1:10,0.1:20,0.01:30,0.005"):
    rules = step_rules.split(",")
    last = float(rules[-1])
    for r in rules[:-1]:
        value, upto = r.split(":")
        if step < int(upto):
            return float(value)
    return last


def lr_multiplier(name, step, num_warmup_steps=0, num_training_steps=None, base_lr=None):
    """lr_lambda(step) of diffusers' get_scheduler(name, ...)."""
    if name not in SCHEDULES:
        raise ValueError(f"{name} is not a valid SchedulerType, please select one of {list(SCHEDULES)}.")
    if name == "constant":
        return _constant(step, 0, 0)
    if name == "piecewise_constant":
        return _piecewise_constant(step)
    if num_warmup_steps is None:
        raise ValueError(f"{name} requires `num_warmup_steps`, please provide that argument.")
    if name == "constant_with_warmup":
        return _constant_with_warmup(step, num_warmup_steps, 0)
    if num_training_steps is None:
        raise ValueError(f"{name} requires `num_training_steps`, please provide that argument.")
    if name == "linear":
        return _linear(step, num_warmup_steps, num_training_steps)
    if name == "cosine":
        return _cosine(step, num_warmup_steps, num_training_steps)
    if name == "cosine_with_restarts":
        return _cosine_with_restarts(step, num_warmup_steps, num_training_steps)
    return _polynomial(step, num_warmup_steps, num_training_steps, base_lr)


def lr_at(name, base_lr, optimizer_step, num_warmup_steps, num_training_steps, num_processes=1):
    """Learning rate of optimiser step `optimizer_step` (0-based) as the reference run would use it."""
    tick = optimizer_step * max(1, num_processes)
    return base_lr * lr_multiplier(name, tick, num_warmup_steps, num_training_steps, base_lr)
