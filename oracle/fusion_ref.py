"""TEST INFRASTRUCTURE ONLY — fp32 functional restatement of the LR2PPO fusion model over a plain
state_dict: Mlp, XiT block, Actor / Critic / Reward forward.  Dropout is the identity (eval mode)
unless masks are supplied.  Validated against the imported reference modules by
oracle/make_golden.py (-> tests/golden/fusion_small.pt) and tests/test_oracle_cpu.py.
"""
import math

import torch
import torch.nn.functional as F


def mlp(sd, pre, x):
    """ref: finetune/ppo.py:154-170 (fc1 -> exact GELU -> fc2, drop p=0)."""
    h = F.gelu(F.linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"]))
    return F.linear(h, sd[pre + "fc2.weight"], sd[pre + "fc2.bias"])


def xit(sd, pre, x, y, heads=8, masks=None):
    """ref: finetune/xit.py:9-148.  x [n,Sq,E] queries, y [n,Skv,E] keys/values.
    softmax over keys FIRST, then divide by sqrt(E) (xit.py:142-143); no mask (xit.py:134-140 drops it).
    masks: optional dict {1,2,3: keep/(1-p) multipliers} for the three dropout sites."""
    E = x.shape[-1]
    a = pre + "0.0.0.fn."
    lx = F.layer_norm(x, (E,), sd[a + "0.ln_x.weight"], sd[a + "0.ln_x.bias"], 1e-5)
    ly = F.layer_norm(y, (E,), sd[a + "0.ln_y.weight"], sd[a + "0.ln_y.bias"], 1e-5)
    q = F.linear(lx, sd[a + "1.queries.weight"], sd[a + "1.queries.bias"])
    k = F.linear(ly, sd[a + "1.keys.weight"], sd[a + "1.keys.bias"])
    v = F.linear(ly, sd[a + "1.values.weight"], sd[a + "1.values.bias"])
    n, Sq, _ = q.shape
    Skv = k.shape[1]
    dh = E // heads
    qh = q.view(n, Sq, heads, dh).transpose(1, 2)
    kh = k.view(n, Skv, heads, dh).transpose(1, 2)
    vh = v.view(n, Skv, heads, dh).transpose(1, 2)
    att = torch.softmax(qh @ kh.transpose(-1, -2), dim=-1) / math.sqrt(E)
    o = (att @ vh).transpose(1, 2).reshape(n, Sq, E)
    o = F.linear(o, sd[a + "1.projection.weight"], sd[a + "1.projection.bias"])
    if masks is not None:
        o = o * masks[1]
    x1 = o + x
    f = pre + "0.0.1.fn."
    l2 = F.layer_norm(x1, (E,), sd[f + "0.weight"], sd[f + "0.bias"], 1e-5)
    h = F.gelu(F.linear(l2, sd[f + "1.0.weight"], sd[f + "1.0.bias"]))
    if masks is not None:
        h = h * masks[2]
    h = F.linear(h, sd[f + "1.3.weight"], sd[f + "1.3.bias"])
    if masks is not None:
        h = h * masks[3]
    x2 = h + x1
    return F.layer_norm(x2, (E,), sd[pre + "1.0.weight"], sd[pre + "1.0.bias"], 1e-5)


def _view_masks(masks, keys, n, S):
    """{site: [n*S, C]} -> {1,2,3: [n, S, C]} for the xit() call that owns sites `keys` (None stays None)."""
    if masks is None:
        return None
    return {i + 1: masks[k].view(n, S, -1) for i, k in enumerate(keys)}


def fusion_body(sd, text, img, masks=None):
    """ref: finetune/ppo.py:214-225 — projections, XiT, concat, out_layer.  text [bs,T,S,E], img [bs,T,I,E].
    masks: train mode — {1,2,3: multiplier [bs*T*S, C]} for the three nn.Dropout(0.1) of `xit` (finetune/xit.py:26-41),
    as produced by oracle/philox.py; None = eval mode."""
    bs, T, S, E = text.shape
    tf = mlp(sd, "text_proj.", text).reshape(bs * T, S, E)
    imf = mlp(sd, "img_proj.", img).reshape(bs * T, -1, E)
    x = xit(sd, "xit.", tf, imf, masks=_view_masks(masks, (1, 2, 3), bs * T, S))
    x = torch.cat([x, imf], dim=1)
    x = mlp(sd, "out_layer.", x.reshape(bs * T, -1))
    return x.view(bs, T, E)


def actor_forward(sd, text, img, masks=None):
    """ref: finetune/ppo.py:214-232 (mode 'reg'): logits [bs*T]."""
    x = fusion_body(sd, text, img, masks)
    return F.linear(x, sd["head.weight"], sd["head.bias"]).view(-1)


def critic_forward(sd, text, img, index, masks=None):
    """ref: finetune/ppo.py:265-297 / :318-350 — gather by index, body, + pos_emb, xitt, head, last token.
    masks (train mode): sites 1-3 for `xit` as in fusion_body, 4-6 [bs*T, C] for `xitt`."""
    bs = text.shape[0]
    bi = torch.arange(bs, device=text.device).view(bs, 1)
    text = text[bi, index]
    img = img[bi, index]
    x = fusion_body(sd, text, img, masks)
    T = x.shape[1]
    x = x + sd["pos_emb.weight"][:T].unsqueeze(0)
    x = xit(sd, "xitt.", x, x, masks=_view_masks(masks, (4, 5, 6), bs, T))
    logits = F.linear(x, sd["head.weight"], sd["head.bias"])
    return logits[:, -1].reshape(bs)


# ---- "trad" (MSLR / MQ2008) variants: finetune/ppo_trad.py:142-283 -------------------------------------------
def trad_body(sd, text):
    """text [bs, T, 768] -> [bs, T, 768]: xit((x, x)) on single-token items, cat with x, out_layer."""
    bs, T, E = text.shape
    x = text.reshape(bs * T, 1, E)
    f = xit(sd, "xit.", x, x)
    f = torch.cat([f, x], dim=1)
    return mlp(sd, "out_layer.", f.reshape(bs * T, -1)).view(bs, T, E)


def trad_actor_forward(sd, text):
    return F.linear(trad_body(sd, text), sd["head.weight"], sd["head.bias"]).view(-1)


def trad_critic_forward(sd, text, index):
    bs = text.shape[0]
    text = text[torch.arange(bs, device=text.device).view(bs, 1), index]
    x = trad_body(sd, text)
    T = x.shape[1]
    x = x + sd["pos_emb.weight"][:T].unsqueeze(0)
    x = xit(sd, "xitt.", x, x)
    return F.linear(x, sd["head.weight"], sd["head.bias"])[:, -1].reshape(bs)
