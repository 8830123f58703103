from diff_pruning_b200.models import UNet2DModel, UNet2DOutput  # noqa: F401  (pickle path diffusers.models.unet_2d.UNet2DModel)
