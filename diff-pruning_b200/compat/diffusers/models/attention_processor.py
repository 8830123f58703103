"""Attention + the processor classes a reference pickle names (attention_processor.py:415, :870).  The engine implements the legacy
AttnProcessor math with the explicit (stale after pruning) `scale` for both — AttnProcessor2_0 itself fails on pruned inner widths."""
from diff_pruning_b200.models import Attention  # noqa: F401


class AttnProcessor:
    pass


class AttnProcessor2_0:
    pass
