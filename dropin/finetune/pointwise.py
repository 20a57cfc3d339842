"""finetune/pointwise.py of the reference tree -- the file `pointwise.sh` launches -- on the B200 path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap  # noqa: E402,F401
from lr2ppo_b200.scripts.pointwise import main  # noqa: E402
from lr2ppo_b200.data import PointwiseClips as MovieNet, get_dataloader  # noqa: E402,F401
from lr2ppo_b200.models import Classifier, Mlp  # noqa: E402,F401
from lr2ppo_b200.stages import build_optimizer, pointwise_evaluate as evaluate  # noqa: E402,F401
from lr2ppo_b200.stages import pointwise_train_model as train_model  # noqa: E402,F401

if __name__ == "__main__":
    main()
