from lr2ppo_b200.tower import PositionwiseFeedForward  # noqa: F401
