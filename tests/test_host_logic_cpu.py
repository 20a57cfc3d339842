"""Host-side logic that needs no GPU: GEMM planning rules (must agree with the kernel selection in
lr2_gemm_bf16), bench.py's shared config object and reference-arm plumbing, the no-fallback guards."""
import importlib
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_gemm_fills_the_cta_pairs():
    from lr2ppo_b200 import ops
    # big activations x weight: already >= 74 pair tiles -> no split
    assert ops.plan_gemm(9408, 3072, 768) == 1
    assert ops.plan_gemm(9408, 768, 3072) == 1
    # weight gradients: 12 x 3 = 36 pair tiles, K = 9408 -> 2 splits (72 of 74 pairs busy)
    assert ops.plan_gemm(3072, 768, 9408) == 2
    assert ops.plan_gemm(768, 3072, 9408) == 2
    # 768 x 768: 9 tiles -> 8 splits
    assert ops.plan_gemm(768, 768, 9408) == 8
    # not eligible for the pair kernel: N % 256, tiny M, short K, transposed output
    assert ops.plan_gemm(9408, 200, 768) == 1
    assert ops.plan_gemm(48, 768, 3072) == 1
    assert ops.plan_gemm(3072, 162816, 48) == 1
    assert ops.plan_gemm(3072, 768, 9408, transposed_out=True) == 1
    # never splits finer than 8 k-blocks per split, never more than 16
    for M, N, K in [(256, 256, 512), (512, 256, 100000), (2048, 2048, 640)]:
        s = ops.plan_gemm(M, N, K)
        assert 1 <= s <= 16 and (s == 1 or (K + 63) // 64 // s >= 8 - 1)


def test_plan_small_gemm():
    from lr2ppo_b200 import ops
    assert ops.plan_small_gemm(48, 768, 3072) == 12          # 6 tiles, 48 k-blocks
    assert ops.plan_small_gemm(96, 3072, 768) == 1           # 12 k-blocks: too short to split
    assert ops.plan_small_gemm(96, 3072, 2048) == 6          # 24 tiles, 32 k-blocks -> 148 // 24
    assert ops.plan_small_gemm(9408, 768, 768) == 1          # many tiles
    assert ops.plan_small_gemm(48, 768, 512) == 1            # short K
    assert ops.plan_small_gemm(48, 770, 3072) == 1           # N % 8 (split-K reduce is 8-wide)


def test_bench_config_is_shared_by_both_arms_and_reference_arm_needs_no_gpu(monkeypatch, capsys):
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    cfg = bench.workload_config(1)
    assert cfg["workload"].startswith("configs[3]") and cfg["queries_per_step_per_gpu"] == 24
    assert bench.workload_config(8)["parallelism"] == "dp8"
    # reference arm: rank != 0 does nothing; rank 0 prints one JSON line built from cpu_stage3's result
    fake = {"value": 3.0, "unit": "queries/s", "cores": 8, "kind": "port", "sample": "s", "ms_per_step": 8000.0,
            "bs": 24}
    monkeypatch.setattr(bench, "cpu_stage3", lambda steps, warmup, budget_s: dict(fake))
    args = bench.argparse.Namespace(gpus=2, steps=3, warmup=1)
    bench.run_reference(args, rank=1)
    assert capsys.readouterr().out == ""
    bench.run_reference(args, rank=0)
    line = json.loads(capsys.readouterr().out)
    assert line["impl"] == "reference" and line["metric"] == "LR2PPO stage-3 train queries/sec"
    assert line["config"]["workload"] == bench.workload_config(2)["workload"]
    assert line["e2e"] == {"value": 3.0, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 8


def test_no_cpu_fallback_guards():
    from lr2ppo_b200 import _lib, ops
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(_lib.Lr2Error):
        ops.gemm(a, a)                                       # CPU tensors are refused, never computed on the host
    from lr2ppo_b200.feed import DeviceFeeder
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            DeviceFeeder((torch.zeros(4),), "cuda:0")


def test_adamw_chunk_span_arithmetic():
    """FusedAdamW launches chunk spans; the two-phase (data-parallel) and row-sharded steps cut intervals out of them.
    Every chunk must be launched exactly once by phase 1 + phase 2, and foreign rows never."""
    from lr2ppo_b200.optim import subtract_span
    assert subtract_span([(0, 10)], (3, 4)) == [(0, 3), (7, 3)]
    assert subtract_span([(0, 10)], (0, 10)) == []
    assert subtract_span([(0, 10)], (0, 4)) == [(4, 6)]
    assert subtract_span([(0, 10)], (6, 4)) == [(0, 6)]
    assert subtract_span([(0, 3), (7, 3)], (2, 6)) == [(0, 2), (8, 2)]
    assert subtract_span([(0, 10)], (12, 5)) == [(0, 10)]
    # a 100-chunk table whose tensor [20, 60) is row-sharded 4 ways (this rank owns the 3rd quarter) and launched early
    spans = subtract_span([(0, 100)], (20, 40)) + [(40, 10)]
    early = (40, 10)
    late = sorted(subtract_span(spans, early))
    launched = sorted([early] + late)
    covered = [c for a, n in launched for c in range(a, a + n)]
    assert covered == list(range(0, 20)) + list(range(40, 50)) + list(range(60, 100))     # once each, foreign rows never


def test_checkpoint_contract_key_names_order_shapes_fp32():
    """state_dict key names / shapes / fp32 are the checkpoint contract (strict=True loads at
    finetune/ppo.py:361,371).  tests/golden_util.param_specs lists the reference's keys (test_oracle_cpu loads them into
    the imported reference modules with strict=True); here the product modules must expose exactly the same keys, in
    the same order, with the same shapes, as fp32 — on CPU, no kernel involved."""
    import argparse
    import torch
    from lr2ppo_b200 import models, ppo
    from tests import golden_util
    cfg = golden_util.FUSION_CFG
    args = argparse.Namespace(mode="reg", labels_num=3, seq_length=cfg["seq_length"], max_imgs=cfg["max_imgs"],
                              visual_feat_dim=cfg["feat"])
    for kind, cls in (("actor", models.Actor), ("reward", models.Reward)):
        m = cls(args, args)
        sd = m.state_dict()
        specs = golden_util.param_specs(kind, cfg)
        assert list(sd.keys()) == [n for n, _ in specs]
        assert [tuple(v.shape) for v in sd.values()] == [tuple(s) for _, s in specs]
        assert all(v.dtype == torch.float32 for v in sd.values())
        assert sd["out_layer.fc1.weight"].shape == (3072, 162816)
        if kind == "actor":
            # stage-3 file = actor.* + critic.* (finetune/ppo.py:912-914 saves the ActorCritic)
            ac = ppo.ActorCritic.__new__(ppo.ActorCritic)
            torch.nn.Module.__init__(ac)
            ac.actor = m
            assert all(k.startswith("actor.") for k in ac.state_dict())
            assert [k[len("actor."):] for k in ac.state_dict()] == list(sd.keys())
        del m, sd


def test_tower_checkpoint_contract_names_and_shapes():
    """The TencentPretrain towers built by lr2ppo_b200.tower.build_model carry the reference's parameter names and
    shapes (tests/golden/tower.pt stores the reference's named_parameters() list of both towers)."""
    import argparse
    import os
    import torch
    from lr2ppo_b200 import tower
    from tests import golden_util
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tower.pt"))
    vit = dict(emb_size=768, feedforward_size=3072, hidden_size=768, hidden_act="gelu", heads_num=12, layers_num=12,
               dropout=0.1, max_seq_length=197, embedding=["patch", "pos"], remove_embedding_layernorm=True,
               encoder="transformer", mask="fully_visible", layernorm_positioning="pre", image_height=224,
               image_width=224, patch_size=16)
    roberta = dict(emb_size=768, feedforward_size=3072, hidden_size=768, hidden_act="gelu", heads_num=12,
                   layers_num=12, max_seq_length=514, dropout=0.1, embedding=["word", "pos", "seg"],
                   encoder="transformer", mask="fully_visible")
    for kind, cfg in (("vit", vit), ("roberta", roberta)):
        model = tower.build_model(argparse.Namespace(**cfg), vocab_size=golden_util.TOWER_VOCAB)
        got = [(n, tuple(p.shape)) for n, p in model.named_parameters()]
        want = [(n, tuple(s)) for n, s in gold[kind]["names"]]
        assert dict(got) == dict(want), set(dict(got)) ^ set(dict(want))
        assert all(p.dtype == torch.float32 for p in model.parameters())


def test_trad_checkpoint_contract():
    """MSLR ("trad") variants: same key contract against the reference's key list (golden_util.trad_param_specs,
    loaded strict=True into the imported reference modules by test_oracle_cpu)."""
    import argparse
    import torch
    from lr2ppo_b200 import trad
    from tests import golden_util
    args = argparse.Namespace(mode="reg", labels_num=5)
    for kind, cls in (("actor", trad.Actor), ("critic", trad.Critic), ("reward", trad.Reward)):
        sd = cls(args, args).state_dict()
        specs = golden_util.trad_param_specs(kind)
        assert list(sd.keys()) == [n for n, _ in specs]
        assert [tuple(v.shape) for v in sd.values()] == [tuple(s) for _, s in specs]
        assert all(v.dtype == torch.float32 for v in sd.values())
