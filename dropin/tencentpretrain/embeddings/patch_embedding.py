from lr2ppo_b200.tower import PatchEmbedding  # noqa: F401
