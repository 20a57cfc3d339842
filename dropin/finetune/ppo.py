"""finetune/ppo.py of the reference tree -- the file `ppo.sh` launches -- on the B200 path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap  # noqa: E402,F401
from lr2ppo_b200.scripts.ppo import main  # noqa: E402
from lr2ppo_b200.data import PpoPairs as MovieNet, get_dataloader  # noqa: E402,F401
from lr2ppo_b200.ppo import (Actor, ActorCritic, Critic, Mlp, RankLoss, Reward, build_optimizer,  # noqa: E402,F401
                             clipped_value_loss, evaluate, log, train_model)

if __name__ == "__main__":
    main()
