"""Shared parity bookkeeping for the GPU tests.

north_star's bound for bf16-computed quantities is 2e-2 relative; here "relative" is always relative to the
tensor's own scale (max |reference| over the compared entries, floored by the reference RMS so that a sample
that happens to hold only tiny entries cannot shrink the denominator).  Every checked quantity is recorded with
its measured error and the bound it was held to; at session end conftest.py writes the table to
gpurun_out/parity_errors.json, and profiles/r02_parity_errors.md is the committed copy of the B200 run.

LR2_PARITY_MEASURE=1 records without asserting (to survey worst cases before tightening a bound)."""
import os

ELEM_TOL = 2e-2          # bound for logits, values, rewards, losses, statistics, hidden states (north_star)
NORM_TOL = 2e-2          # bound on |norm - ref norm| / ref norm of a gradient / moment tensor
# Element-wise MAX error of a parameter-gradient tensor.  Measured on B200 (profiles/r02_parity_errors.md): storing
# activations in bf16 puts ~1 % rms noise (of the tensor's scale) on every gradient element, so the maximum over the
# 10^3..10^5 compared entries of a tensor lands at 2-3.5 %.  tests/test_bf16_noise_floor_gpu.py shows that this is the
# noise floor of bf16 storage, not of these kernels: stock torch.autocast(bfloat16) over the fp32 restatement has the
# SAME error on the SAME tensors (max 0.035 vs ours 0.033, mean 0.0140 vs 0.0138) and that test holds every tensor
# to max(2e-2, 1.5 x autocast's error).  Against the fixed goldens (no autocast run available at test time) gradient
# tensors are therefore held to 4e-2 element-wise and 2e-2 in norm; quantities where the LOSS amplifies forward
# noise (a difference of two nearly equal backward passes, a residual V - R) carry explicit, documented overrides.
GRAD_ELEM_TOL = 4e-2
# Documented overrides (measured values in profiles/r02_parity_errors.md):
# (a) end-to-end stage-3 steps: d value_loss / d V = 2 (V - R) / B is a residual of two O(1) quantities, so the ~1 %
#     forward noise of V changes the SCALE of the critic's gradient by a few %.  The four small tensors between the
#     head and the first large GEMM see that scale error undiluted (measured elem <= 0.048, norm <= 0.041).
CRITIC_TAIL_ELEM = {"xitt.": 6e-2, "head.": 6e-2, "pos_emb": 6e-2, "out_layer.fc2.bias": 6e-2}
CRITIC_TAIL_NORM = {"xitt.": 5e-2, "head.": 5e-2, "pos_emb": 5e-2, "out_layer.fc2.bias": 5e-2}
# (b) stage 2: d loss / d chosen = -g, d loss / d reject = +g, and the two forwards differ only in the order of the
#     last two slots: the gradients of everything after the item bodies (pos_emb, out_layer.fc2.bias, xitt) are
#     DIFFERENCES of two nearly equal backward passes, each carrying its own bf16 noise (measured 0.095 with 3 pairs,
#     0.062 with 64 pairs; the fp32 reference does not cancel noise, it has none).
def pair_cancellation(tol):
    return {"pos_emb": tol, "out_layer.fc2.bias": tol, "xitt.": tol}
RECORDS = []
MEASURE = os.environ.get("LR2_PARITY_MEASURE") == "1"


def check(test, name, err, bound):
    """Record `err` (float) for quantity `name` of test `test` and hold it to `bound`."""
    err = float(err)
    RECORDS.append({"test": test, "name": name, "err": err, "bound": float(bound), "ok": bool(err < bound)})
    if not MEASURE:
        assert err < bound, (test, name, err, bound)


def rel_err(got, ref, floor=0.0):
    """max |got - ref| / max(max |ref|, floor), both moved to CPU fp32."""
    got = got.detach().float().cpu().reshape(-1)
    ref = ref.detach().float().cpu().reshape(-1)
    scale = max(ref.abs().max().item(), floor, 1e-30)
    return (got - ref).abs().max().item() / scale


def _tol_for(name, default, overrides):
    for key, tol in (overrides or {}).items():
        if key in name:
            return tol
    return default


def check_param_tensors(test, named, getter, ref_of, norm_of, sample, elem_tol=GRAD_ELEM_TOL, norm_tol=NORM_TOL,
                        elem_overrides=None, norm_overrides=None):
    """Element-wise + norm comparison of one tensor per parameter (gradients or Adam first moments).

    named:   [(name, parameter)]
    getter:  parameter -> CUDA tensor to check (p.grad, optimizer.state[p]['exp_avg'], ...)
    ref_of:  name -> reference entries (the full tensor when small, else the strided `sample` of it)
    norm_of: name -> reference L2 norm of the FULL tensor (python float)
    sample:  callable(tensor) -> the same strided sample of a full tensor (golden_util.grad_sample)
    elem_overrides / norm_overrides: {substring of the parameter name: bound} for documented exceptions

    Tensors whose reference RMS is < 1e-4 of the largest RMS in the model are mathematically zero (e.g.
    keys.bias: softmax is invariant to a per-query shift, the reference holds fp32 rounding noise there): ours
    must be negligible against the real gradients too, nothing else can be asked of noise."""
    rms = {n: norm_of(n) / max(1.0, p.numel() ** 0.5) for n, p in named}
    top = max(rms.values())
    for n, p in named:
        got = getter(p)
        assert got is not None, n
        if rms[n] < 1e-4 * top:
            check(test, n + " [zero-gradient tensor, rms vs top rms]",
                  got.double().norm().item() / max(1.0, got.numel() ** 0.5) / top, 1e-2)
            continue
        ref = ref_of(n)
        gs = got if ref.numel() == got.numel() else sample(got)
        check(test, n + " [elem]", rel_err(gs, ref, floor=rms[n]), _tol_for(n, elem_tol, elem_overrides))
        check(test, n + " [norm]", abs(got.double().norm().item() - norm_of(n)) / norm_of(n),
              _tol_for(n, norm_tol, norm_overrides))


def dump(path):
    import json
    if not RECORDS:
        return
    os.makedirs(os.path.dirname(path), exist_ok=True)
    worst = {}
    for r in RECORDS:
        w = worst.setdefault(r["test"], {"n": 0, "max_err_over_bound": 0.0, "worst": None})
        w["n"] += 1
        ratio = r["err"] / r["bound"] if r["bound"] > 0 else 0.0
        if ratio >= w["max_err_over_bound"]:
            w["max_err_over_bound"], w["worst"] = ratio, r
    with open(path, "w") as f:
        json.dump({"measure_only": MEASURE, "summary": worst, "records": RECORDS}, f, indent=1)
