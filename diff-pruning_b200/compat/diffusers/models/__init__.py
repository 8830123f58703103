from diff_pruning_b200.models import UNet2DModel  # noqa: F401
from . import attention_processor, embeddings, resnet, unet_2d, unet_2d_blocks  # noqa: F401
