"""tencentpretrain/model_builder.py: `build_model(args)` -> Embedding + TransformerEncoder towers (ViT-B/16,
RoBERTa-base) on the tcgen05 kernels."""
from lr2ppo_b200.tower import build_model  # noqa: F401
