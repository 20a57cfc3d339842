from lr2ppo_b200.tower import TransformerLayer  # noqa: F401
