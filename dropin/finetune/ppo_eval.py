"""finetune/ppo_eval.py of the reference tree -- the file `ppo_eval.sh` launches -- on the B200 path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap  # noqa: E402,F401
from lr2ppo_b200.scripts.ppo_eval import main  # noqa: E402
from lr2ppo_b200.data import EvalClips as MovieNet, get_dataloader  # noqa: E402,F401
from lr2ppo_b200.ppo import Actor, ActorCritic, Critic, Mlp, RankLoss, Reward, evaluate  # noqa: E402,F401

if __name__ == "__main__":
    main()
