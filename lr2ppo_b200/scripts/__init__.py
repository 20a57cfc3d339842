"""main() of the four multimodal stage scripts (finetune/{pointwise,reward_pair_dataloader,ppo,ppo_eval}.py) on the
B200 path.  `dropin/finetune/*.py` are the files the reference's `.sh` launchers name; they re-export these."""
