from lr2ppo_b200.tower import LayerNorm  # noqa: F401
