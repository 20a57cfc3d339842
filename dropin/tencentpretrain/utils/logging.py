from lr2ppo_b200.runtime import init_logger  # noqa: F401
