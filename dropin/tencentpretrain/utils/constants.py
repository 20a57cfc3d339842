from lr2ppo_b200.tokenizers import CLS_TOKEN, MASK_TOKEN, PAD_TOKEN, SEP_TOKEN, UNK_TOKEN  # noqa: F401
