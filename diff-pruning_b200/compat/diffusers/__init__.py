"""`import diffusers` surface of the hot path, backed by diff_pruning_b200 (see ../README.md).

Names the reference's scripts import: ddpm_prune.py:1-2, ddpm_train.py:19-23, ddpm_sample.py:1.  Sub-modules mirror the paths a
whole-module pickle written by the reference refers to (`torch.save(model)` at ddpm_prune.py:135 stores classes by module path:
diffusers.models.unet_2d.UNet2DModel, ...unet_2d_blocks.*, ...resnet.*, ...attention_processor.*, ...embeddings.*,
diffusers.configuration_utils.FrozenDict), so `torch.load(args.pruned_model_ckpt)` at ddpm_train.py:292 / ddpm_sample.py:27 resolves
them to the engine-backed classes (UNet2DModel.__setstate__ adopts the reference's attribute layout).
"""
from diff_pruning_b200.models import DDPMScheduler, UNet2DModel, UNet2DOutput  # noqa: F401
from diff_pruning_b200.sampling import DDIMPipeline, DDIMScheduler, DDPMPipeline, DiffusionPipeline  # noqa: F401

from . import configuration_utils, models, optimization, training_utils, utils  # noqa: F401,E402

__version__ = "0.17.0.dev0+dpb200"

# torch >= 2.6 unpickles with weights_only=True by default; the scripts load WHOLE modules with a bare torch.load(path)
# (ddpm_train.py:292, ddpm_sample.py:27), which only works when every class in the pickle is allow-listed.
from diff_pruning_b200.checkpoint import allow_module_pickles as _allow  # noqa: E402

_allow()
