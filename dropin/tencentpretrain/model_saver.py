"""tencentpretrain/model_saver.py: `save_model(model, path)` -- same file format (a plain fp32 state_dict), written
asynchronously; `wait()` blocks until the file is on disk."""
from lr2ppo_b200.checkpoint import save_model, wait  # noqa: F401
