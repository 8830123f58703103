"""`import accelerate` surface used by the reference's scripts (ddpm_train.py:255-261,277-284,348-361,386,407,420,455-477,536;
ddpm_sample.py:22,45-79; `import accelerate` only at ddpm_prune.py:12), as a thin layer over torch.distributed — SURVEY.md §8(b1).

One process per GPU (torchrun / torch.distributed.launch --use_env sets RANK / LOCAL_RANK / WORLD_SIZE); the NCCL process group
is created on first use.  Data parallelism follows torch DDP semantics as accelerate applies them: every process gets its own
share of the batches (`prepare(dataloader)`), gradients are AVERAGED over processes after `backward()`; here that is ONE
all-reduce of the engine's flat gradient arena (Parameter.grad are views into it) instead of DDP's buckets.
`mixed_precision="bf16"` selects the engine's bf16 tensor tier for the UNet (what torch.autocast(bfloat16) does to conv / linear
in ddpm_train.py:255-261); fp16 (which needs a GradScaler in the reference) is not offered.
"""
from __future__ import annotations

import contextlib
import os

import torch

from . import logging, utils  # noqa: F401
from .utils import ProjectConfiguration  # noqa: F401

__version__ = "0.20.0+dpb200"


class _State:
    def __init__(self, acc):
        self.acc = acc

    def __repr__(self):
        a = self.acc
        return (f"Distributed environment: {'MULTI_GPU' if a.num_processes > 1 else 'NO'}\nNum processes: {a.num_processes}\n"
                f"Process index: {a.process_index}\nLocal process index: {a.local_process_index}\nDevice: {a.device}\n"
                f"Mixed precision type: {a.mixed_precision}\n")


class _ShardedLoader:
    """What `accelerator.prepare(dataloader)` returns: this process's share of the batches (batch k goes to process k mod world,
    all processes see the same number of batches), already on the accelerator's device."""

    def __init__(self, loader, acc):
        self.loader, self.acc = loader, acc
        self.dataset = getattr(loader, "dataset", None)
        self.batch_size = getattr(loader, "batch_size", None)

    def __len__(self):
        return len(self.loader) // self.acc.num_processes if self.acc.num_processes > 1 else len(self.loader)

    def _to_dev(self, b):
        if torch.is_tensor(b):
            return b.to(self.acc.device, non_blocking=True)
        if isinstance(b, (list, tuple)):
            return type(b)(self._to_dev(x) for x in b)
        if isinstance(b, dict):
            return {k: self._to_dev(v) for k, v in b.items()}
        return b

    def __iter__(self):
        world, rank, n = self.acc.num_processes, self.acc.process_index, len(self)
        taken = 0
        for k, batch in enumerate(self.loader):
            if world > 1 and k % world != rank:
                continue
            if taken >= n:
                break
            taken += 1
            yield self._to_dev(batch)


class Accelerator:
    def __init__(self, gradient_accumulation_steps: int = 1, mixed_precision=None, log_with=None, project_dir=None,
                 project_config=None, cpu: bool = False, **unused):
        if mixed_precision in ("fp16",):
            raise NotImplementedError("diff_pruning_b200: mixed_precision='fp16' (GradScaler path) is not provided; use 'no' or 'bf16'")
        self.mixed_precision = mixed_precision or "no"
        self.gradient_accumulation_steps = int(gradient_accumulation_steps)
        self.log_with, self.project_dir, self.project_config = log_with, project_dir, project_config
        self.num_processes = int(os.environ.get("WORLD_SIZE", "1"))
        self.process_index = int(os.environ.get("RANK", "0"))
        self.local_process_index = int(os.environ.get("LOCAL_RANK", "0"))
        use_cuda = torch.cuda.is_available() and not cpu
        self.device = torch.device("cuda", self.local_process_index) if use_cuda else torch.device("cpu")
        if use_cuda:
            torch.cuda.set_device(self.device)
        if self.num_processes > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                os.environ.setdefault("MASTER_PORT", "29500")
                if use_cuda:
                    dist.init_process_group("nccl", device_id=self.device)
                else:
                    dist.init_process_group("gloo")
        self.state = _State(self)
        self.sync_gradients = True
        self._accum_count = 0
        self._models = []
        self._trackers = {}

    # ---- process topology
    @property
    def is_main_process(self):
        return self.process_index == 0

    @property
    def is_local_main_process(self):
        return self.local_process_index == 0

    def wait_for_everyone(self):
        if self.num_processes > 1:
            import torch.distributed as dist
            dist.barrier()

    def print(self, *a, **kw):
        if self.is_local_main_process:
            print(*a, **kw)

    # ---- prepare / unwrap
    def prepare(self, *objs):
        out = []
        for o in objs:
            if isinstance(o, torch.nn.Module):
                o = o.to(self.device)
                if self.mixed_precision == "bf16":
                    o.__dict__["_dpb200_compute"] = "bf16"     # read by diff_pruning_b200.engine.unet_apply
                self._models.append(o)
            elif isinstance(o, torch.utils.data.DataLoader):
                o = _ShardedLoader(o, self)
            out.append(o)
        return out[0] if len(out) == 1 else tuple(out)

    def unwrap_model(self, model, **unused):
        return model

    # ---- the step
    @contextlib.contextmanager
    def accumulate(self, model=None):
        self._accum_count += 1
        self.sync_gradients = self._accum_count % self.gradient_accumulation_steps == 0
        yield

    def backward(self, loss, **kw):
        if self.gradient_accumulation_steps > 1:
            loss = loss / self.gradient_accumulation_steps
        loss.backward(**kw)
        if self.num_processes > 1 and self.sync_gradients:
            self._average_gradients()

    def _average_gradients(self):
        import torch.distributed as dist
        for m in self._models:
            arena = None
            for plan in m.__dict__.get("_dpb200_plans", {}).values():
                if plan.need_grad and all(p.grad is not None and p.grad.data_ptr() == plan._grad_views[id(p)].data_ptr() for p in plan.params):
                    arena = plan.grad_arena
                    break
            if arena is not None:                       # every Parameter.grad is a view of this flat buffer: one collective
                dist.all_reduce(arena, op=dist.ReduceOp.SUM)
                arena.div_(self.num_processes)
            else:
                for p in m.parameters():
                    if p.grad is not None:
                        dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
                        p.grad.div_(self.num_processes)

    def clip_grad_norm_(self, parameters, max_norm, norm_type=2):
        return torch.nn.utils.clip_grad_norm_(parameters, max_norm, norm_type=norm_type)

    # ---- trackers (tensorboard only when asked for and importable; otherwise logs are dropped like accelerate does with log_with=None)
    def init_trackers(self, project_name, config=None, **unused):
        if self.log_with == "tensorboard" and self.is_main_process:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self._trackers["tensorboard"] = SummaryWriter(os.path.join(self.project_dir or ".", project_name))
            except Exception:
                pass

    def get_tracker(self, name, unwrap=False):
        return self._trackers[name]

    def log(self, values: dict, step=None):
        tb = self._trackers.get("tensorboard")
        if tb is not None:
            for k, v in values.items():
                if isinstance(v, (int, float)):
                    tb.add_scalar(k, v, global_step=step)

    def end_training(self):
        for t in self._trackers.values():
            try:
                t.close()
            except Exception:
                pass
        self._trackers.clear()
