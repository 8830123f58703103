"""Learning-rate multipliers of `diffusers.optimization.get_scheduler` (optimization.py:40-78, 123-183, 282-340) as plain
functions of the optimisation step, for the finetune loop of ddpm_train.py:340-346,464: the reference builds a LambdaLR around
these; FinetuneStepper reads `stepper.lr` at every step, so the caller sets

    stepper.lr = base_lr * lr_multiplier(args.lr_scheduler, global_step, args.lr_warmup_steps, max_train_steps)

before `stepper.step(...)` (global_step counts completed optimiser steps, i.e. LambdaLR's `last_epoch`).  The reference's default
is "constant".  Host-only."""
from __future__ import annotations

import math


def lr_multiplier(name: str, step: int, num_warmup_steps: int = 0, num_training_steps: int = 0, num_cycles: float = 0.5) -> float:
    if name == "constant":
        return 1.0
    if name == "constant_with_warmup":
        return step / max(1.0, float(num_warmup_steps)) if step < num_warmup_steps else 1.0
    if name not in ("linear", "cosine"):
        raise NotImplementedError(f"lr scheduler {name!r} (supported: constant, constant_with_warmup, linear, cosine)")
    if step < num_warmup_steps:
        return step / float(max(1, num_warmup_steps))
    span = float(max(1, num_training_steps - num_warmup_steps))
    if name == "linear":
        return max(0.0, (num_training_steps - step) / span)
    progress = (step - num_warmup_steps) / span
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))
