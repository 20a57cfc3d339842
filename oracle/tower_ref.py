"""TEST INFRASTRUCTURE ONLY — fp32 functional restatement of the TencentPretrain towers over a state_dict with
the reference's key names (embedding.*, encoder.*).  Follows tencentpretrain/embeddings/*.py,
encoders/transformer_encoder.py:48-138, layers/transformer.py:50-73, layers/multi_headed_attn.py:27-76,
layers/position_ffn.py:12-15, layers/layer_norm.py:16-21.  Dropout = identity (eval).  Validated against the
reference's build_model output by tests/golden/tower.pt (oracle/make_golden.py tower)."""
import math

import torch
import torch.nn.functional as F

from .restate import tencent_layernorm


def _ln(sd, pre, x):
    return tencent_layernorm(x, sd[pre + "gamma"], sd[pre + "beta"], 1e-6)


def embed_patch(sd, img, patch=16):
    B = img.shape[0]
    w = sd["embedding.patch.projection.weight"]
    x = F.conv2d(img, w, stride=patch).flatten(2).transpose(1, 2)
    x = torch.cat([sd["embedding.patch.cls_emb"].expand(B, -1, -1), x], dim=1)
    return x + sd["embedding.pos.embedding.weight"][: x.shape[1]].unsqueeze(0)


def embed_word(sd, src, seg):
    S = src.shape[1]
    x = sd["embedding.word.embedding.weight"][src] + sd["embedding.pos.embedding.weight"][:S].unsqueeze(0)
    if "embedding.seg.embedding.weight" in sd:
        x = x + sd["embedding.seg.embedding.weight"][seg]
    return _ln(sd, "embedding.layer_norm.", x)


def attention(sd, pre, x, mask, heads):
    B, S, E = x.shape
    dh = E // heads
    q, k, v = (F.linear(x, sd[f"{pre}linear_layers.{i}.weight"], sd[f"{pre}linear_layers.{i}.bias"])
               .view(B, S, heads, dh).transpose(1, 2) for i in range(3))
    scores = q @ k.transpose(-2, -1) / math.sqrt(float(dh)) + mask
    out = (torch.softmax(scores, dim=-1) @ v).transpose(1, 2).reshape(B, S, E)
    return F.linear(out, sd[pre + "final_linear.weight"], sd[pre + "final_linear.bias"])


def ffn(sd, pre, x):
    return F.linear(F.gelu(F.linear(x, sd[pre + "linear_1.weight"], sd[pre + "linear_1.bias"])),
                    sd[pre + "linear_2.weight"], sd[pre + "linear_2.bias"])


def encoder(sd, emb, seg, layers, heads, pre_ln):
    B, S, _ = emb.shape
    mask = (1.0 - (seg > 0).float().unsqueeze(1).repeat(1, S, 1).unsqueeze(1)) * -10000.0
    h = emb
    for i in range(layers):
        p = f"encoder.transformer.{i}."
        if pre_ln:
            h = h + attention(sd, p + "self_attn.", _ln(sd, p + "layer_norm_1.", h), mask, heads)
            h = ffn(sd, p + "feed_forward.", _ln(sd, p + "layer_norm_2.", h)) + h
        else:
            inter = _ln(sd, p + "layer_norm_1.", attention(sd, p + "self_attn.", h, mask, heads) + h)
            h = _ln(sd, p + "layer_norm_2.", ffn(sd, p + "feed_forward.", inter) + inter)
    return _ln(sd, "encoder.layer_norm.", h) if pre_ln else h
