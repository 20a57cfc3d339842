"""finetune/misc.py of the reference tree (`from misc import *`): distributed start-up and rank helpers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap  # noqa: E402,F401
from lr2ppo_b200.runtime import (get_rank, get_world_size, init_distributed_mode,  # noqa: E402,F401
                                 is_dist_avail_and_initialized, is_main_process, mkdir, setup_for_distributed,
                                 setup_seed)
