"""tencentpretrain/model_loader.py: `load_model(model, path)` (strict=False, like the reference)."""
import torch


def load_model(model, model_path):
    target = model.module if hasattr(model, "module") else model
    target.load_state_dict(torch.load(model_path, map_location="cpu"), strict=False)
    return model
