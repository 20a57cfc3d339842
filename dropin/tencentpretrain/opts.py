"""tencentpretrain/opts.py: the flag groups the stage scripts register (`finetune_opts`, `tokenizer_opts`, `adv_opts`)."""
from lr2ppo_b200.cli import (adv_opts, finetune_opts, log_opts, model_opts, optimization_opts,  # noqa: F401
                             tokenizer_opts, training_opts)
