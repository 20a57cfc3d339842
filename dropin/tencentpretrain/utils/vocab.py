from lr2ppo_b200.tokenizers import Vocab  # noqa: F401
