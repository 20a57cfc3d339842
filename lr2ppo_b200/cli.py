"""Command-line contract of the stage scripts: the flag groups of tencentpretrain/opts.py (finetune_opts :129,
tokenizer_opts :175, adv_opts :222 and the groups they pull in) and the config merge of
tencentpretrain/utils/config.py:6-23, so that the flag arrays of pointwise.sh, reward_pair_dataloader.sh, ppo.sh and
ppo_eval.sh parse unchanged.

The flags are declared as data (one row per flag) and registered by one loop; names, types, defaults and choices are
the contract, the rest is ours.  Flags that select components outside the LR2PPO hot path (recurrent / CNN encoders,
speech front-end, adversarial training ...) are still accepted -- the scripts pass some of them -- and rejected later,
where the component would be built, with a message that says so."""
import argparse
import json
import sys

TOKENIZERS = ["bert", "bpe", "char", "space", "xlmroberta", "image", "text_image", "virtual"]
EMBEDDINGS = ["word", "pos", "seg", "sinusoidalpos", "patch", "speech", "word_patch", "dual"]
ENCODERS = ["transformer", "rnn", "lstm", "gru", "birnn", "bilstm", "bigru", "gatedcnn", "dual"]
LEVELS = ["ERROR", "INFO", "DEBUG", "NOTSET"]
SCHEDULERS = ["linear", "cosine", "cosine_with_restarts", "polynomial", "constant", "constant_with_warmup",
              "inverse_sqrt", "tri_stage"]

# (flag, kind, default, extra)   kind: a type, "flag" (store_true) or ("choice", [...]); extra: argparse keywords
_ROWS = {
    "paths": [
        ("pretrained_model_path", str, None, {}), ("output_model_path", str, "models/finetuned_model.bin", {}),
        ("train_path", str, None, {}), ("dev_path", str, None, {}), ("test_path", str, None, {}),
        ("config_path", str, "models/bert/base_config.json", {}),
    ],
    "model": [
        ("embedding", ("choice", EMBEDDINGS), "word", {"nargs": "+"}),
        ("tgt_embedding", ("choice", EMBEDDINGS), "word", {"nargs": "+"}),
        ("max_seq_length", int, 512, {}), ("relative_position_embedding", "flag", None, {}),
        ("share_embedding", "flag", None, {}), ("remove_embedding_layernorm", "flag", None, {}),
        ("factorized_embedding_parameterization", "flag", None, {}),
        ("encoder", ("choice", ENCODERS), "transformer", {}), ("decoder", ("choice", [None, "transformer"]), None, {}),
        ("mask", ("choice", ["fully_visible", "causal", "causal_with_prefix"]), "fully_visible", {}),
        ("layernorm_positioning", ("choice", ["pre", "post"]), "post", {}),
        ("feed_forward", ("choice", ["dense", "gated"]), "dense", {}),
        ("relative_attention_buckets_num", int, 32, {}), ("remove_attention_scale", "flag", None, {}),
        ("remove_transformer_bias", "flag", None, {}), ("layernorm", ("choice", ["normal", "t5"]), "normal", {}),
        ("bidirectional", "flag", None, {}), ("parameter_sharing", "flag", None, {}),
        ("has_residual_attention", "flag", None, {}), ("has_lmtarget_bias", "flag", None, {}),
        ("target", ("choice", ["sp", "lm", "mlm", "bilm", "cls", "clr"]), "mlm", {"nargs": "+"}),
        ("tie_weights", "flag", None, {}), ("pooling", ("choice", ["mean", "max", "first", "last"]), "first", {}),
    ],
    "vision": [
        ("image_height", int, 256, {}), ("image_width", int, 256, {}), ("patch_size", int, 16, {}),
        ("channels_num", int, 3, {}), ("image_preprocess", str, ["crop", "normalize"], {"nargs": "+"}),
    ],
    "audio": [
        ("sampling_rate", int, 16000, {}),
        ("audio_preprocess", str, ["normalize_means", "normalize_vars", "ceptral_normalize"], {"nargs": "+"}),
        ("max_audio_frames", int, 6000, {}), ("conv_layers_num", int, 2, {}), ("audio_feature_size", int, 80, {}),
        ("conv_channels", int, 1024, {}), ("conv_kernel_sizes", int, [5, 5], {"nargs": "+"}),
    ],
    "optimization": [
        ("learning_rate", float, 2e-5, {}), ("warmup", float, 0.1, {}), ("decay", float, 0.5, {}),
        ("fp16", "flag", None, {}), ("fp16_opt_level", ("choice", ["O0", "O1", "O2", "O3"]), "O1", {}),
        ("optimizer", ("choice", ["adamw", "adafactor"]), "adamw", {}),
        ("scheduler", ("choice", SCHEDULERS), "linear", {}),
    ],
    "training": [
        ("batch_size", int, 32, {}), ("seq_length", int, 128, {}), ("max_imgs", int, 32, {}),
        ("visual_feat_dim", int, -1, {}), ("dropout", float, 0.1, {}), ("epochs_num", int, 3, {}),
        ("report_steps", int, 100, {}), ("seed", int, 7, {}),
    ],
    "log": [
        ("log_path", str, None, {}), ("log_level", ("choice", LEVELS), "INFO", {}),
        ("log_file_level", ("choice", LEVELS), "INFO", {}),
    ],
    "tokenizer": [
        ("tokenizer", ("choice", TOKENIZERS), "bert", {}), ("vocab_path", str, None, {}),
        ("merges_path", str, None, {}), ("spm_model_path", str, None, {}),
        ("do_lower_case", ("choice", ["true", "false"]), "true", {}), ("vqgan_model_path", str, None, {}),
        ("vqgan_config_path", str, None, {}),
    ],
    "adv": [
        ("use_adv", "flag", None, {}), ("adv_type", ("choice", ["fgm", "pgd"]), "fgm", {}),
        ("fgm_epsilon", float, 1e-6, {}), ("pgd_k", int, 3, {}), ("pgd_epsilon", float, 1.0, {}),
        ("pgd_alpha", float, 0.3, {}),
    ],
    # flags every multimodal stage script adds itself (finetune/ppo.py:713-732, finetune/pointwise.py:443-462)
    "stage_common": [
        ("mode", str, "reg", {}), ("vit_pretrained_model_path", str, None, {}),
        ("vit_config_path", str, "models/bert/base_config.json", {}),
        ("vit_tokenizer", ("choice", TOKENIZERS), None, {}), ("vit_encoder", ("choice", ENCODERS), None, {}),
        ("dist_url", str, "env://", {}), ("max_tags", int, 32, {}), ("exp_name", str, None, {}),
        ("use_pairwise", "flag", None, {}),
    ],
    "stage12": [("soft_targets", "flag", None, {}), ("soft_alpha", float, 0.5, {})],
    "ppo": [
        ("reward_model_path", str, None, {}), ("max_timesteps", int, 5, {}), ("update_timesteps", int, 300, {}),
        ("eps_clip", float, 0.2, {}), ("kl_div_loss_weight", float, 0.1, {}), ("entropy_weight", float, 0.1, {}),
        ("value_clip", float, 0.4, {}), ("critic_learning_rate", float, 2e-6, {}),
    ],
}


def add_group(parser, group):
    for name, kind, default, extra in _ROWS[group]:
        if kind == "flag":
            parser.add_argument("--" + name, action="store_true")
        elif isinstance(kind, tuple):
            parser.add_argument("--" + name, choices=kind[1], default=default, **extra)
        else:
            parser.add_argument("--" + name, type=kind, default=default, **extra)


def model_opts(parser):
    for g in ("model", "vision", "audio"):
        add_group(parser, g)


def log_opts(parser):
    add_group(parser, "log")


def optimization_opts(parser):
    add_group(parser, "optimization")


def training_opts(parser):
    add_group(parser, "training")
    log_opts(parser)


def finetune_opts(parser):
    add_group(parser, "paths")
    model_opts(parser)
    optimization_opts(parser)
    training_opts(parser)


def tokenizer_opts(parser):
    add_group(parser, "tokenizer")


def adv_opts(parser):
    add_group(parser, "adv")


def stage_parser(stage):
    """The parser each stage script builds in its main(): stage in 'pointwise' | 'reward_pair_dataloader' | 'ppo' |
    'ppo_eval'."""
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    finetune_opts(parser)
    tokenizer_opts(parser)
    if stage in ("pointwise", "reward_pair_dataloader"):
        add_group(parser, "stage12")
    adv_opts(parser)
    add_group(parser, "stage_common")
    if stage in ("ppo", "ppo_eval"):
        add_group(parser, "ppo")
    return parser


def vit_namespace(args):
    """Second namespace in which every `vit_xxx` flag also appears as `xxx` (finetune/ppo.py:735-744)."""
    d = dict(vars(args))
    for k, v in vars(args).items():
        if "vit_" in k:
            d[k[4:]] = v
    return argparse.Namespace(**d)


def load_hyperparam(default_args, argv=None):
    """defaults < JSON config file < flags given on the command line (tencentpretrain/utils/config.py:6-23; a flag
    counts as given when `--name` appears in argv)."""
    argv = sys.argv if argv is None else argv
    with open(default_args.config_path, mode="r", encoding="utf-8") as f:
        from_file = json.load(f)
    merged = dict(vars(default_args))
    given = {a[2:]: merged[a[2:]] for a in argv if a.startswith("--") and "local_rank" not in a and a[2:] in merged}
    merged.update(from_file)
    merged.update(given)
    return argparse.Namespace(**merged)
