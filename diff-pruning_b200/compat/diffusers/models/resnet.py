from diff_pruning_b200.models import Downsample2D, ResnetBlock2D, Upsample2D  # noqa: F401  (ddpm_prune.py:112)
