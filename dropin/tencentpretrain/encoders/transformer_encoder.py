from lr2ppo_b200.tower import TransformerEncoder  # noqa: F401
