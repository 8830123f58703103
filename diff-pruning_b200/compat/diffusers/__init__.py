"""`import diffusers` surface of the hot path, backed by diff_pruning_b200 (see ../README.md)."""
from diff_pruning_b200.models import DDPMScheduler, UNet2DModel, UNet2DOutput  # noqa: F401
from diff_pruning_b200.sampling import DDIMPipeline, DDIMScheduler, DDPMPipeline  # noqa: F401

__version__ = "0.17.0.dev0+dpb200"
