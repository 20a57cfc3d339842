"""ctypes binding of liblr2ppo_b200.so (the C ABI declared in include/lr2ppo_b200.h).

This is the whole "FFI" of the package: torch is used only to own device memory and
streams; every kernel is reached through the plain-pointer entry points below.
There is no CPU fallback: if the shared library is missing, or a call returns a
negative code, an exception is raised.
"""
import ctypes
import os
from ctypes import c_int, c_longlong, c_float, c_void_p, c_ulonglong, c_uint, c_char_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblr2ppo_b200.so")

_lib = None

vp, i32, i64, f32, u64, u32 = c_void_p, c_int, c_longlong, c_float, c_ulonglong, c_uint

# name -> (restype, argtypes); mirrors include/lr2ppo_b200.h one to one
SIGNATURES = {
    "lr2_abi_version": (i32, []),
    "lr2_last_error_string": (c_char_p, [i32]),
    "lr2_check_device": (i32, []),
    "lr2_launch_count": (i64, []),
    "lr2_note_launches": (None, [i32]),
    "lr2_gemm_workspace_bytes": (i64, [i32, i32, i32, i32, i64]),
    "lr2_gemm_bf16": (i32, [vp, i64, i32, vp, i64, i32, vp, i64, i32, i32, i32, i32, i32, i32, vp, vp, i64, vp,
                            f32, f32, u64, u32, vp, i32, vp, i32, vp]),
    "lr2_gemm_wgrad_adamw": (i32, [vp, i64, vp, i64, i32, i32, i32, vp, vp, vp, vp, vp, f32, vp]),
    "lr2_layernorm_fwd": (i32, [vp, vp, vp, vp, vp, i64, i32, f32, i32, i32, i32, i32, vp]),
    "lr2_layernorm_bwd_partials_floats": (i64, [i32]),
    "lr2_layernorm_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, f32, i32, i32, i32, i32, f32, u64,
                                u32, vp, vp]),
    "lr2_bump_counter": (i32, [vp, u64, vp]),
    "lr2_xattn_fwd": (i32, [vp, i64, vp, vp, i64, vp, i64, i32, i32, i32, i32, i32, f32, f32, vp]),
    "lr2_xattn_bwd": (i32, [vp, i64, vp, vp, i64, vp, i64, vp, i64, vp, vp, i64, i32, i32, i32, i32, i32, f32, f32,
                            vp]),
    "lr2_mha_fwd": (i32, [vp, vp, vp, i64, vp, vp, i64, vp, i32, i32, i32, i32, f32, f32, u64, vp, vp]),
    "lr2_mha_bwd": (i32, [vp, vp, vp, i64, vp, vp, vp, i64, vp, vp, vp, vp, i64, i32, i32, i32, i32, f32, f32, u64, vp,
                          vp]),
    "lr2_embed_sum": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, i32, vp]),
    "lr2_embed_scatter_add": (i32, [vp, vp, vp, i64, i32, vp]),
    "lr2_patchify": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "lr2_dropout_bf16": (i32, [vp, vp, i64, f32, u64, u32, vp, vp]),
    "lr2_bias_gelu_rows": (i32, [vp, vp, vp, vp, i64, i32, vp]),
    "lr2_cast_gather_bf16": (i32, [vp, vp, vp, i32, i32, i32, i64, vp]),
    "lr2_gather_rows_bf16": (i32, [vp, vp, vp, i32, i32, i32, i64, vp]),
    "lr2_rows_copy_bf16": (i32, [vp, i64, i64, vp, i64, i64, i64, i64, i32, i32, vp]),
    "lr2_colsum_partials_floats": (i64, [i32]),
    "lr2_colsum_bf16": (i32, [vp, i64, i64, i32, vp, vp, i32, vp]),
    "lr2_rowdot_fwd": (i32, [vp, i64, i64, vp, vp, vp, i32, i32, vp]),
    "lr2_rowdot_bwd": (i32, [vp, i64, i64, vp, vp, vp, vp, vp, i32, i32, vp]),
    "lr2_add_pos_fwd": (i32, [vp, vp, i32, i32, i32, vp]),
    "lr2_add_pos_bwd": (i32, [vp, vp, i32, i32, i32, vp]),
    "lr2_cast_f32_to_bf16": (i32, [vp, vp, i64, vp]),
    "lr2_cast_bf16_to_f32": (i32, [vp, vp, i64, vp]),
    "lr2_ppo_policy_loss": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, f32, f32, f32, f32, vp, vp, vp, vp, vp, vp, vp]),
    "lr2_clipped_value_loss": (i32, [vp, vp, vp, i32, f32, vp, vp, vp]),
    "lr2_pair_hinge_loss": (i32, [vp, vp, i32, f32, vp, vp, vp, vp]),
    "lr2_smooth_l1_loss": (i32, [vp, vp, i64, f32, vp, vp, vp]),
    "lr2_ppo_rollout": (i32, [vp, vp, i32, i32, i32, vp, vp, vp]),
    "lr2_rank_sample": (i32, [vp, vp, i32, i32, i32, vp, vp, vp]),
    "lr2_gae_scan": (i32, [vp, vp, vp, i32, i32, f32, f32, vp, vp, vp]),
    "lr2_rank_logprob": (i32, [vp, vp, i32, i32, vp, vp, vp, vp, vp]),
    "lr2_ppo_clip_surrogate": (i32, [vp, vp, vp, i32, f32, i32, f32, vp, vp, vp, vp]),
    "lr2_ndcg_at_k": (i32, [vp, vp, vp, i32, i32, i64, vp, i32, vp, vp, vp, vp]),
    "lr2_ndcg_presorted": (i32, [vp, vp, vp, i32, i32, vp, i32, vp, vp, vp, vp]),
    "lr2_adamw_chunk_elems": (i32, []),
    "lr2_adamw_multi": (i32, [vp, vp, vp, i64, vp, vp]),
}


class Lr2Error(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Lr2Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI drifted from the header
        fn.restype = res
        fn.argtypes = args
    if lib.lr2_abi_version() != 1:
        raise Lr2Error("ABI version mismatch")
    _lib = lib
    return lib


def check(code, what=""):
    if code != 0:
        msg = load().lr2_last_error_string(code).decode()
        raise Lr2Error(f"{what}: error {code} ({msg})")


# When PROFILE is a list, every C-ABI call is bracketed by CUDA events on the launching stream and
# (name, args, start, end) is appended (bench.py's per-kernel roofline pass). None = no overhead.
PROFILE = None


def run(fn, *args):
    """Call a C-ABI entry point, raise on a non-zero return code."""
    if PROFILE is None:
        code = fn(*args)
    else:
        import torch
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        code = fn(*args)
        e1.record()
        PROFILE.append((fn.__name__, args, e0, e1))
    if code != 0:
        check(code, fn.__name__)


def launch_count():
    return load().lr2_launch_count()


def ptr(t):
    """Device (or host) address of a tensor, or None."""
    return None if t is None else t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
