from lr2ppo_b200.tower import PosEmbedding  # noqa: F401
