from lr2ppo_b200.optim import *  # noqa: F401,F403
from lr2ppo_b200.optim import AdamW, str2optimizer, str2scheduler  # noqa: F401
