"""The structural group derivation (pruning.build_groups) against the groups torch_pruning's DependencyGraph produced in
the reference run (tests/golden/cifar_cfg1.pt), through the whole interactive prune sequence.  CPU only, no compute."""
import torch

from conftest import expand, load_golden
import diff_pruning_b200 as dp
from diff_pruning_b200 import pruning


def norm(items):
    return sorted((n, k, tuple(sorted(i))) for n, k, i in items)


def test_groups_match_reference_through_the_prune_sequence():
    G = load_golden("cifar_cfg1.pt")["variants"]["taylor"]
    torch.manual_seed(0)
    m = dp.UNet2DModel(**dp.CIFAR10_DDPM_CONFIG)
    morder = pruning.traced_module_order(m)
    first = pruning.build_groups(m, ignored_layers=[m.conv_out], module_order=morder)
    assert [g["root"] for g in first] == [g["root"] for g in G["groups"]]          # same 50 groups, same order
    mods = dict(m.named_modules())
    for ref in G["groups"]:
        mine = next(g for g in pruning.build_groups(m, ignored_layers=[m.conv_out], module_order=morder) if g["root"] == ref["root"])
        assert mine["channels"] == ref["channels"] and mine["ch_groups"] == ref["ch_groups"], ref["root"]
        assert norm(mine["items"]) == norm((n, k, expand(i)) for n, k, i in ref["items"]), ref["root"]
        pruning.apply_group(mods, mine["items"], ref["idxs"], mine["channels"])     # the reference's selection
    pruning.fix_static_attributes(m)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == G["pruned_shapes"]
    assert sum(p.numel() for p in m.parameters()) == 19851157
    # the pruned module tree still builds a plan-able / traceable network
    with torch.no_grad(), dp.trace_mode():
        out = m(torch.randn(1, 3, 32, 32), torch.ones(1).long()).sample
    assert out.shape == (1, 3, 32, 32)


def test_groups_lsun_family_counts():
    torch.manual_seed(0)
    m = dp.UNet2DModel(**dict(dp.LSUN256_DDPM_CONFIG, block_out_channels=(32, 32, 64, 64, 128, 128)))
    gs = pruning.build_groups(m, ignored_layers=[m.conv_out])
    prod = {n for g in gs for n, k, _ in g["items"] if k == "out"}
    layers = {n for n, mod in m.named_modules() if isinstance(mod, (torch.nn.Conv2d, torch.nn.Linear))}
    assert prod == layers - {"conv_out"}          # every conv / linear output is in exactly one group


def test_lsun_architecture_magnitude_prune_sequence_matches_reference():
    """BASELINE config 3's six-level architecture (reduced widths), `--pruner magnitude`, ratio 0.05, through the compat call
    sequence of ddpm_prune.py:79-116: group order / members / ch_groups, the channels the reference selected in every group,
    post-prune shapes and the op / parameter counts (tests/golden/lsun_struct_magnitude.pt, tools/gen_golden.py `lsun_struct`)."""
    import os
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "diff-pruning_b200", "compat"))
    import torch_pruning as tp
    from diffusers.models.resnet import Downsample2D, Upsample2D
    G = load_golden("lsun_struct_magnitude.pt")
    torch.manual_seed(0)
    m = dp.UNet2DModel(**G["cfg"]).eval()
    ex = {"sample": torch.randn(1, 3, 64, 64), "timestep": torch.ones((1,)).long()}
    macs, params = tp.utils.count_ops_and_params(m, ex)
    assert [macs, params] == G["base"]
    pr = tp.pruner.MagnitudePruner(m, ex, importance=tp.importance.MagnitudeImportance(), iterative_steps=1, channel_groups={},
                                   ch_sparsity=G["ratio"], ignored_layers=[m.conv_out])
    seen = []
    for g, ref in zip(pr.step(interactive=True), G["groups"]):
        assert g.root == ref["root"] and g.channels == ref["channels"], (g.root, ref["root"])
        assert norm(g.items) == norm((n, k, expand(i)) for n, k, i in ref["items"]), ref["root"]
        assert sorted(g.idxs) == sorted(ref["idxs"]), ref["root"]            # the same channels go
        seen.append(g.root)
        g.prune()
    assert seen == [g["root"] for g in G["groups"]]
    for mod in m.modules():
        if isinstance(mod, (Upsample2D, Downsample2D)):
            mod.channels = mod.conv.in_channels
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == G["pruned_shapes"]
    assert list(tp.utils.count_ops_and_params(m, ex)) == G["pruned"]
