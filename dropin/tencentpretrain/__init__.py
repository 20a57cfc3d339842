"""`tencentpretrain` namespace of the reference tree, re-exporting the B200 implementations (lr2ppo_b200) under the
module paths the stage scripts import from.  Components of other model families (recurrent / CNN encoders, speech
front-end, pre-training targets, corpus builders) are not part of the LR2PPO hot path and are not provided."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap  # noqa: E402,F401
