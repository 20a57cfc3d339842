"""finetune/reward_pair_dataloader.py of the reference tree -- the file `reward_pair_dataloader.sh` launches -- on the B200 path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap  # noqa: E402,F401
from lr2ppo_b200.scripts.reward_pair_dataloader import main  # noqa: E402
from lr2ppo_b200.data import RewardPairs as MovieNet, get_dataloader  # noqa: E402,F401
from lr2ppo_b200.models import Mlp, PairClassifier as Classifier  # noqa: E402,F401
from lr2ppo_b200.stages import build_optimizer, reward_evaluate as evaluate  # noqa: E402,F401
from lr2ppo_b200.stages import reward_train_model as train_model  # noqa: E402,F401

if __name__ == "__main__":
    main()
