"""Applying channel pruning to the module tree (host side, integer/index work).

Replaces the slicing done by torch_pruning's layer pruners — ddpm_exp/torch_pruning/pruner/function.py:85-146
(ConvPruner), :168-207 (LinearPruner), :274-302 (GroupNormPruner): keep-index selection on weight / bias and on the
accumulated ``.grad`` (so later groups of the same prune pass score the already-sliced tensors, exactly like the
reference's interactive loop ddpm_prune.py:108-109), plus the static-attribute fix of ddpm_prune.py:112-116.
Slicing runs on whatever device the parameters live on (torch.index_select on the GPU for the product path).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

import torch
import torch.nn as nn

from .models import Downsample2D, Upsample2D


def _keep(n: int, drop: Sequence[int], device) -> torch.Tensor:
    mask = torch.ones(n, dtype=torch.bool)
    mask[torch.as_tensor(sorted(set(int(i) for i in drop)), dtype=torch.long)] = False
    return torch.nonzero(mask).flatten().to(device)


def _slice_param(p: nn.Parameter, keep: torch.Tensor, dim: int) -> nn.Parameter:
    new = nn.Parameter(torch.index_select(p.data, dim, keep).contiguous(), requires_grad=p.requires_grad)
    if p.grad is not None:
        new.grad = torch.index_select(p.grad, dim, keep).contiguous()
    return new


def prune_out_channels(layer: nn.Module, drop: Sequence[int]) -> None:
    """Conv2d / Linear output channels or GroupNorm channels (function.py:88-105, 171-185, 277-294)."""
    if isinstance(layer, nn.GroupNorm):
        keep = _keep(layer.num_channels, drop, layer.weight.device)
        layer.num_channels = int(keep.numel())          # num_groups is NOT touched (function.py:274-294)
        layer.weight = _slice_param(layer.weight, keep, 0)
        layer.bias = _slice_param(layer.bias, keep, 0)
        return
    n = layer.weight.shape[0]
    keep = _keep(n, drop, layer.weight.device)
    layer.weight = _slice_param(layer.weight, keep, 0)
    if layer.bias is not None:
        layer.bias = _slice_param(layer.bias, keep, 0)
    if isinstance(layer, nn.Conv2d):
        layer.out_channels = int(keep.numel())
    else:
        layer.out_features = int(keep.numel())


def prune_in_channels(layer: nn.Module, drop: Sequence[int]) -> None:
    """Conv2d / Linear input channels (function.py:107-126, 187-199); bias untouched."""
    n = layer.weight.shape[1]
    keep = _keep(n, drop, layer.weight.device)
    layer.weight = _slice_param(layer.weight, keep, 1)
    if isinstance(layer, nn.Conv2d):
        layer.in_channels = int(keep.numel())
    else:
        layer.in_features = int(keep.numel())


def apply_group(modules: Dict[str, nn.Module], items: Iterable[Tuple[str, str, Sequence[int]]], selected: Sequence[int],
                channels: int) -> None:
    """Prune one group: `items` are (layer_name, kind in {out,in,gn}, full index list) where index lists map the root's
    channels positionally (a layer fed by both halves of a concat carries the two maps merged: len = parts*channels);
    `selected` are root-channel positions to remove."""
    for name, kind, idxs in items:
        idxs = list(idxs)
        parts = len(idxs) // channels
        assert parts * channels == len(idxs), (name, len(idxs), channels)
        drop = [idxs[q * channels + j] for q in range(parts) for j in selected]
        layer = modules[name]
        if kind == "in":
            prune_in_channels(layer, drop)
        else:
            prune_out_channels(layer, drop)


def fix_static_attributes(model: nn.Module) -> None:
    """ddpm_prune.py:112-116."""
    for m in model.modules():
        if isinstance(m, (Upsample2D, Downsample2D)):
            m.channels = m.conv.in_channels
            m.out_channels = m.conv.out_channels
