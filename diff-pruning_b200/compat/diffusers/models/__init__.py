from diff_pruning_b200.models import UNet2DModel  # noqa: F401
from . import resnet  # noqa: F401
