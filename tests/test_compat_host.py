"""The import-compatible `diffusers` / `torch_pruning` surface (compat/), CPU structure checks."""
import os
import sys

import torch

from conftest import ROOT, load_golden

sys.path.insert(0, os.path.join(ROOT, "diff-pruning_b200", "compat"))


def test_compat_names_and_counts():
    import diffusers
    import torch_pruning as tp
    from diffusers.models.resnet import Downsample2D, Upsample2D   # ddpm_prune.py:112
    import diff_pruning_b200 as dp
    assert diffusers.UNet2DModel is dp.UNet2DModel and Downsample2D is not None and Upsample2D is not None
    torch.manual_seed(0)
    m = diffusers.UNet2DModel(**dp.CIFAR10_DDPM_CONFIG).eval()
    ex = {"sample": torch.randn(1, 3, 32, 32), "timestep": torch.ones((1,)).long()}
    macs, params = tp.utils.count_ops_and_params(m, ex)
    G = load_golden("cifar_cfg1.pt")["variants"]["vendored"]
    assert params == G["base"][1] == 35746307
    assert macs == G["base"][0]                                   # same counter convention as the reference (6.064 G)
    # magnitude pruning through the unchanged call sequence of ddpm_prune.py:79-116 reaches the published architecture size
    pr = tp.pruner.MagnitudePruner(m, ex, importance=tp.importance.MagnitudeImportance(), iterative_steps=1, channel_groups={},
                                   ch_sparsity=0.3, ignored_layers=[m.conv_out])
    for g in pr.step(interactive=True):
        g.prune()
    for mod in m.modules():
        if isinstance(mod, (Upsample2D, Downsample2D)):
            mod.channels = mod.conv.in_channels
    macs2, params2 = tp.utils.count_ops_and_params(m, ex)
    assert params2 == G["pruned"][1] == 19851157 and macs2 == G["pruned"][0]    # 3.392 G (assets/exp.png)
