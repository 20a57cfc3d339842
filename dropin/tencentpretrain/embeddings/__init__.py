from lr2ppo_b200.tower import (Embedding, PatchEmbedding, PosEmbedding, SegEmbedding, WordEmbedding,  # noqa: F401
                               str2embedding)

__all__ = ["Embedding", "WordEmbedding", "PosEmbedding", "SegEmbedding", "PatchEmbedding", "str2embedding"]
