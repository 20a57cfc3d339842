from lr2ppo_b200.tower import LayerNorm, MultiHeadedAttention, PositionwiseFeedForward, TransformerLayer  # noqa: F401
