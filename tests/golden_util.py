"""Seeded tensors shared by oracle/make_golden.py (generator) and the parity tests (consumers).
Weights are never stored: both sides regenerate them on the CPU from the same seeds."""
import torch

FUSION_CFG = dict(seq_length=196, max_imgs=16, feat=768, bs=2, tags=2)
SEEDS = dict(actor=101, critic=202, reward=303)
INPUT_SEED = 9001


def init_params_(model, seed):
    """Deterministic init in named_parameters() order: matrices N(0, 0.02) as in finetune/ppo.py:363-365,
    LayerNorm scales 1 + 0.1 N(0,1), other vectors 0.02 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            r = torch.randn(p.shape, generator=g)
            if p.dim() >= 2:
                p.copy_(r * 0.02)
            elif name.endswith("weight") and ("ln_" in name or ".fn.0." in name or name.startswith(("xit.1", "xitt.1"))):
                p.copy_(1.0 + 0.1 * r)
            else:
                p.copy_(r * 0.02)


def make_inputs(kind, cfg=FUSION_CFG):
    g = torch.Generator().manual_seed(INPUT_SEED + SEEDS[kind])
    bs, T, S, I, E = cfg["bs"], cfg["tags"], cfg["seq_length"], cfg["max_imgs"], cfg["feat"]
    text = torch.randn(bs, T, S, E, generator=g)
    img = torch.randn(bs, T, I, E, generator=g)
    tgts = torch.randint(0, 3, (bs, T), generator=g)
    if kind == "actor":
        index = None
    elif kind == "critic":
        index = torch.stack([torch.randperm(T, generator=g) for _ in range(bs)])
    else:
        perm = torch.stack([torch.randperm(T, generator=g) for _ in range(bs)])
        index = torch.cat([torch.arange(T).unsqueeze(0).repeat(bs, 1), perm], dim=1)
    return text, img, tgts, index


def out_grad(kind, n):
    g = torch.Generator().manual_seed(SEEDS[kind] + 5)
    return torch.randn(n, generator=g)


def grad_sample(gr, k=4096):
    flat = gr.reshape(-1)
    step = max(1, flat.numel() // k)
    return flat[::step][:k]


def _xit_specs(pre, E=768):
    H = 4 * E
    a = pre + ".0.0.0.fn."
    f = pre + ".0.0.1.fn."
    out = []
    for ln in ("ln_x", "ln_y"):
        out += [(a + f"0.{ln}.weight", (E,)), (a + f"0.{ln}.bias", (E,))]
    for lin in ("keys", "queries", "values", "projection"):
        out += [(a + f"1.{lin}.weight", (E, E)), (a + f"1.{lin}.bias", (E,))]
    out += [(f + "0.weight", (E,)), (f + "0.bias", (E,)), (f + "1.0.weight", (H, E)), (f + "1.0.bias", (H,)),
            (f + "1.3.weight", (E, H)), (f + "1.3.bias", (E,)), (pre + ".1.0.weight", (E,)), (pre + ".1.0.bias", (E,))]
    return out


def _mlp_specs(pre, i, h, o):
    return [(pre + ".fc1.weight", (h, i)), (pre + ".fc1.bias", (h,)), (pre + ".fc2.weight", (o, h)),
            (pre + ".fc2.bias", (o,))]


def param_specs(kind, cfg=FUSION_CFG):
    """(name, shape) in the reference's named_parameters() order (finetune/ppo.py:196-212, 247-263)."""
    E = cfg["feat"]
    specs = _mlp_specs("text_proj", E, 4 * E, E) + _mlp_specs("img_proj", E, 4 * E, E)
    if kind != "actor":
        specs += [("pos_emb.weight", (4, E))]
    specs += _xit_specs("xit", E)
    if kind != "actor":
        specs += _xit_specs("xitt", E)
    specs += _mlp_specs("out_layer", (cfg["seq_length"] + cfg["max_imgs"]) * E, 4 * E, E)
    specs += [("head.weight", (1, E)), ("head.bias", (1,))]
    return specs


def _is_ln_scale(name):
    return name.endswith("weight") and ("ln_" in name or ".fn.0." in name or name.startswith(("xit.1", "xitt.1")))


def make_state_dict(kind, cfg=FUSION_CFG):
    """Same draws as init_params_ on the reference module, without needing the reference."""
    g = torch.Generator().manual_seed(SEEDS[kind])
    sd = {}
    for name, shape in param_specs(kind, cfg):
        r = torch.randn(shape, generator=g)
        if len(shape) >= 2:
            sd[name] = r * 0.02
        elif _is_ln_scale(name):
            sd[name] = 1.0 + 0.1 * r
        else:
            sd[name] = r * 0.02
    return sd


# ---- "trad" models (finetune/ppo_trad.py:142-283) --------------------------------------------------------
TRAD_SEEDS = dict(actor=411, critic=422, reward=433)


def trad_param_specs(kind, E=768):
    specs = []
    if kind != "actor":
        specs += [("pos_emb.weight", (4, E))]
    specs += _xit_specs("xit", E)
    if kind != "actor":
        specs += _xit_specs("xitt", E)
    specs += _mlp_specs("out_layer", 2 * E, 4 * E, E)
    specs += [("head.weight", (1, E)), ("head.bias", (1,))]
    return specs


def make_trad_state_dict(kind):
    g = torch.Generator().manual_seed(TRAD_SEEDS[kind])
    sd = {}
    for name, shape in trad_param_specs(kind):
        r = torch.randn(shape, generator=g)
        sd[name] = r * 0.02 if len(shape) >= 2 else (1.0 + 0.1 * r if _is_ln_scale(name) else r * 0.02)
    return sd


def trad_inputs(kind, bs=24, docs=2):
    g = torch.Generator().manual_seed(7000 + TRAD_SEEDS[kind])
    text = torch.randn(bs, docs, 768, generator=g)
    tgts = torch.randint(0, 5, (bs, docs), generator=g)
    if kind == "actor":
        index = None
    elif kind == "critic":
        index = torch.stack([torch.randperm(docs, generator=g) for _ in range(bs)])
    else:
        perm = torch.stack([torch.randperm(docs, generator=g) for _ in range(bs)])
        index = torch.cat([torch.arange(docs).unsqueeze(0).repeat(bs, 1), perm], dim=1)
    return text, tgts, index


# ---- stage 1 / stage 2 single training step on the full-size fusion models --------------------------------
STEP_LR = 2e-6


def stage_inputs(stage):
    g = torch.Generator().manual_seed(8100 + stage)
    bs, T = (2, 3) if stage == 1 else (3, 2)
    text = torch.randn(bs, T, 196, 768, generator=g)
    img = torch.randn(bs, 1, 16, 768, generator=g).repeat(1, T, 1, 1)
    tgts = torch.randint(0, 3, (bs, T), generator=g)
    chosen = reject = None
    if stage == 2:
        pick = torch.randint(0, 2, (bs,), generator=g)
        chosen = torch.tensor([[0, 1, 0, 1], [1, 0, 0, 1]])[pick]       # reward_pair_dataloader.py:128-141
        reject = torch.tensor([[0, 1, 1, 0], [1, 0, 1, 0]])[pick]
    return text, img, tgts, chosen, reject


# ---- TencentPretrain towers (build_model with the ViT-B/16 and xlm-roberta base configs) ------------------
TOWER_VOCAB = 50265
TOWER_SEEDS = dict(vit=511, roberta=522)


def make_tower_state_dict(names, seed):
    """names: [(param name, shape)] in the reference's named_parameters() order (stored in the golden file)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in names:
        r = torch.randn(shape, generator=g)
        if name.endswith("gamma"):
            sd[name] = 1.0 + 0.1 * r
        elif len(shape) >= 2:
            sd[name] = r * 0.02
        else:
            sd[name] = r * 0.02
    return sd


def tower_inputs(kind):
    g = torch.Generator().manual_seed(7700 + TOWER_SEEDS[kind])
    if kind == "vit":
        src = torch.randn(2, 3, 224, 224, generator=g)
        seg = torch.ones(2, 197, dtype=torch.int64)
    else:
        src = torch.randint(5, TOWER_VOCAB, (3, 64), generator=g)
        seg = torch.ones(3, 64, dtype=torch.int64)
        seg[1, 40:] = 0                      # padded tail: masked keys (transformer_encoder.py:62-68)
        seg[2, 56:] = 0
    return src, seg


# ---- API-parity modules (VideoTransformer, ProjectionLayer) ---------------------------------------------
VIDEO_CFG = dict(frame_size=16, emb_size=768, layers=2, heads=12, output_dim=512)
API_SEEDS = dict(video=611, proj=622)


def make_api_state_dict(names, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in names:
        r = torch.randn(shape, generator=g)
        is_ln_w = name.endswith("weight") and ("ln_" in name or "layer_norm" in name)
        sd[name] = 1.0 + 0.1 * r if is_ln_w else (r * 0.03 if len(shape) >= 2 else r * 0.02)
    return sd


def api_input(kind):
    g = torch.Generator().manual_seed(7900 + API_SEEDS[kind])
    return torch.randn(5, 16, 768, generator=g) if kind == "video" else torch.randn(7, 33, 768, generator=g)


# ---- round 2: train-mode (replayed dropout masks) and BASELINE-size goldens (oracle/make_golden_r2.py) -----------
TRAIN_SEEDS = dict(actor=70001, critic=70002, reward=70003)     # engine.FusionEngine.dropout_seed of fusion_train.pt
S3 = dict(bs=24, tags=2, seed=4242, actor_seed=81001, critic_seed=81002, lr=1e-3, train_steps=100, warmup=0.1)
S12 = {1: dict(bs=2, tags=20, seed=5151, mask_seed=82001), 2: dict(bs=64, tags=2, seed=5252, mask_seed=82002)}


def stage3_inputs():
    """ppo.sh:21 batch: text [24,2,196,768], img [24,16,768] as the loader yields it (finetune/ppo.py:831), tgts."""
    g = torch.Generator().manual_seed(S3["seed"])
    text = torch.randn(S3["bs"], S3["tags"], 196, 768, generator=g)
    img = torch.randn(S3["bs"], 16, 768, generator=g)
    tgts = torch.randint(0, 3, (S3["bs"], S3["tags"]), generator=g)
    return text, img, tgts


def stage12_inputs(stage):
    """pointwise.sh (2 clips x 20 tags) / reward_pair_dataloader.sh (64 pairs, chosen / reject 4-slot sequences)."""
    c = S12[stage]
    g = torch.Generator().manual_seed(c["seed"])
    text = torch.randn(c["bs"], c["tags"], 196, 768, generator=g)
    img = torch.randn(c["bs"], 16, 768, generator=g)
    tgts = torch.randint(0, 3, (c["bs"], c["tags"]), generator=g)
    chosen = reject = None
    if stage == 2:
        # finetune/reward_pair_dataloader.py:128-141: [0,1,0,1] / [0,1,1,0] or [1,0,0,1] / [1,0,1,0]
        flip = torch.rand(c["bs"], generator=g) < 0.5
        a = torch.tensor([[0, 1, 0, 1], [1, 0, 0, 1]])
        b = torch.tensor([[0, 1, 1, 0], [1, 0, 1, 0]])
        chosen = a[flip.long()]
        reject = b[flip.long()]
    return text, img, tgts, chosen, reject
