from lr2ppo_b200.tower import MultiHeadedAttention  # noqa: F401
