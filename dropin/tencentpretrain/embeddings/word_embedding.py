from lr2ppo_b200.tower import WordEmbedding  # noqa: F401
