from lr2ppo_b200.cli import load_hyperparam  # noqa: F401
