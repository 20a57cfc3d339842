from lr2ppo_b200.tower import TransformerEncoder, str2encoder  # noqa: F401

__all__ = ["TransformerEncoder", "str2encoder"]
