from lr2ppo_b200.runtime import set_seed  # noqa: F401
