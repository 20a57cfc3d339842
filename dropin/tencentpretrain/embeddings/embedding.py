from lr2ppo_b200.tower import Embedding  # noqa: F401
