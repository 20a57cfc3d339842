from lr2ppo_b200.tokenizers import *  # noqa: F401,F403
