from lr2ppo_b200.tower import SegEmbedding  # noqa: F401
