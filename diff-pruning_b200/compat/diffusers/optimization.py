"""diffusers.optimization.get_scheduler (optimization.py:282-340) for ddpm_train.py:340-346: a torch LambdaLR around
diff_pruning_b200.schedules.lr_multiplier (values pinned on tests/golden/lr_schedules.json)."""
from torch.optim.lr_scheduler import LambdaLR

from diff_pruning_b200.schedules import lr_multiplier

_NEEDS_WARMUP = ("constant_with_warmup", "linear", "cosine")
_NEEDS_TOTAL = ("linear", "cosine")


def get_scheduler(name, optimizer, num_warmup_steps=None, num_training_steps=None, num_cycles=1, power=1.0, last_epoch=-1):
    name = getattr(name, "value", name)
    if name in _NEEDS_WARMUP and num_warmup_steps is None:
        raise ValueError(f"{name} requires `num_warmup_steps`, please provide that argument.")
    if name in _NEEDS_TOTAL and num_training_steps is None:
        raise ValueError(f"{name} requires `num_training_steps`, please provide that argument.")
    warm, total = int(num_warmup_steps or 0), int(num_training_steps or 0)
    lr_multiplier(name, 0, warm, total)      # raises NotImplementedError for schedules outside the four the scripts use
    return LambdaLR(optimizer, lambda step: lr_multiplier(name, step, warm, total), last_epoch)
