from diff_pruning_b200.models import AttnDownBlock2D, AttnUpBlock2D, DownBlock2D, UNetMidBlock2D, UpBlock2D  # noqa: F401
