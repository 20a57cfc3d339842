"""ndcg.py of the reference tree: `from ndcg import AverageNDCGMeter` (finetune/ppo.py:31)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _bootstrap  # noqa: E402,F401
from lr2ppo_b200.ndcg import AverageNDCGMeter  # noqa: E402,F401
