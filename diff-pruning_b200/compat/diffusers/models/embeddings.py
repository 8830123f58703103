from diff_pruning_b200.models import TimestepEmbedding, Timesteps  # noqa: F401
