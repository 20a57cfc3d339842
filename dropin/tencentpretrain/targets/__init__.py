"""The fine-tuning scripts never build a pre-training target; `build_model` returns towers
without one."""
str2target = {}
__all__ = ["str2target"]
