"""diffusers.training_utils.EMAModel as the reference MODIFIED it (training_utils.py:182-218): constant decay (`decay` itself, the
warm-up schedule of get_decay is bypassed at :201) and an out-of-place update s = (1-d)*p + d*s (:216).  Used by
ddpm_train.py:320-328 (ctor), :350 (`to`), :388-401 / :489-512 (store / copy_to / restore around sampling and saving), :469 (step)."""
from typing import Iterable

import torch


class EMAModel:
    def __init__(self, parameters: Iterable[torch.nn.Parameter], decay: float = 0.9999, min_decay: float = 0.0,
                 update_after_step: int = 0, use_ema_warmup: bool = False, inv_gamma: float = 1.0, power: float = 2 / 3,
                 model_cls=None, model_config=None, **unused):
        if isinstance(parameters, torch.nn.Module):
            parameters = parameters.parameters()
        self.shadow_params = [p.clone().detach() for p in parameters]
        self.temp_stored_params = None
        self.decay, self.min_decay, self.update_after_step = decay, min_decay, update_after_step
        self.use_ema_warmup, self.inv_gamma, self.power = use_ema_warmup, inv_gamma, power
        self.optimization_step = 0
        self.cur_decay_value = None
        self.model_cls, self.model_config = model_cls, model_config

    def get_decay(self, optimization_step: int) -> float:
        step = max(0, optimization_step - self.update_after_step - 1)
        if step <= 0:
            return 0.0
        v = 1 - (1 + step / self.inv_gamma) ** -self.power if self.use_ema_warmup else (1 + step) / (10 + step)
        return max(min(v, self.decay), self.min_decay)

    @torch.no_grad()
    def step(self, parameters):
        if isinstance(parameters, torch.nn.Module):
            parameters = parameters.parameters()
        self.optimization_step += 1
        decay = self.decay                                   # :201 — the schedule is not consulted
        self.cur_decay_value = decay
        for s, p in zip(self.shadow_params, list(parameters)):
            if p.requires_grad:
                s.data = (1 - decay) * p.data + decay * s.data   # :216
            else:
                s.copy_(p)

    def _wrote(self, parameters):
        # `param.data.copy_` leaves no autograd trace: tell the engine its packed weight copies are stale
        from diff_pruning_b200 import engine
        for p in parameters:
            owner = getattr(p, "_dpb200_owner", None)
            if owner is not None:
                engine.invalidate_packs(owner())
                return

    def copy_to(self, parameters) -> None:
        parameters = list(parameters)
        for s, p in zip(self.shadow_params, parameters):
            p.data.copy_(s.to(p.device).data)

    def to(self, device=None, dtype=None) -> None:
        self.shadow_params = [p.to(device=device, dtype=dtype) if p.is_floating_point() else p.to(device=device)
                              for p in self.shadow_params]

    def store(self, parameters) -> None:
        self.temp_stored_params = [p.detach().cpu().clone() for p in parameters]

    def restore(self, parameters) -> None:
        if self.temp_stored_params is None:
            raise RuntimeError("This ExponentialMovingAverage has no `store()`ed weights to `restore()`")
        for c, p in zip(self.temp_stored_params, parameters):
            p.data.copy_(c.data)
        self.temp_stored_params = None

    def state_dict(self) -> dict:
        return {"decay": self.decay, "min_decay": self.min_decay, "optimization_step": self.optimization_step,
                "update_after_step": self.update_after_step, "use_ema_warmup": self.use_ema_warmup, "inv_gamma": self.inv_gamma,
                "power": self.power, "shadow_params": self.shadow_params}

    def load_state_dict(self, state_dict: dict) -> None:
        for k in ("decay", "min_decay", "optimization_step", "update_after_step", "use_ema_warmup", "inv_gamma", "power"):
            if k in state_dict:
                setattr(self, k, state_dict[k])
        if not 0.0 <= self.decay <= 1.0:
            raise ValueError("Decay must be between 0 and 1")
        sp = state_dict.get("shadow_params")
        if sp is not None:
            if not isinstance(sp, list) or not all(isinstance(p, torch.Tensor) for p in sp):
                raise ValueError("shadow_params must be a list of Tensors")
            self.shadow_params = [p.clone() for p in sp]
