from lr2ppo_b200.tower import Model  # noqa: F401
