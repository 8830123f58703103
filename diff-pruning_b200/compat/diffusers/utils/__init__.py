"""diffusers.utils names used at ddpm_train.py:23,279-283,407: availability probes and the logging verbosity switches."""
import importlib.util
import operator

from packaging import version as _v

from . import logging  # noqa: F401


def is_tensorboard_available() -> bool:
    return importlib.util.find_spec("tensorboard") is not None or importlib.util.find_spec("tensorboardX") is not None


def is_wandb_available() -> bool:
    return importlib.util.find_spec("wandb") is not None


def is_accelerate_available() -> bool:
    return importlib.util.find_spec("accelerate") is not None


def is_accelerate_version(operation: str, version: str) -> bool:
    import accelerate
    ops = {">": operator.gt, ">=": operator.ge, "==": operator.eq, "!=": operator.ne, "<=": operator.le, "<": operator.lt}
    if operation not in ops:
        raise ValueError(f"`operation` must be one of {list(ops)}, received {operation}")
    return ops[operation](_v.parse(_v.parse(accelerate.__version__).base_version), _v.parse(version))
