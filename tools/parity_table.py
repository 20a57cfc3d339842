"""gpurun_out/parity_errors.json (written by a `pytest -m gpu` session, tests/parity.py) -> markdown table.
    python tools/parity_table.py gpurun_out/parity_errors.json > profiles/r02_parity_errors.md"""
import collections
import json
import sys


def main():
    d = json.load(open(sys.argv[1]))
    recs = d["records"]
    by = collections.OrderedDict()
    for r in recs:
        by.setdefault(r["test"], []).append(r)
    print("# Measured parity errors, B200 (`pytest -m gpu`; every quantity the tests hold to a bound)\n")
    print("error = max |ours - reference| / scale of the reference tensor (tests/parity.py); bound = what the test "
          "asserts.\nReference = goldens produced by running the reference's own code (oracle/make_golden*.py) or the "
          "CPU oracle.\n")
    print(f"{len(recs)} quantities in {len(by)} test groups; {sum(not r['ok'] for r in recs)} over their bound.\n")
    print("| test group | quantities | worst error | its bound | worst quantity | max error among [norm] | among forward / loss values |")
    print("|---|---:|---:|---:|---|---:|---:|")
    for t, rs in by.items():
        w = max(rs, key=lambda r: r["err"] / r["bound"] if r["bound"] else 0)
        norms = [r["err"] for r in rs if "[norm]" in r["name"]]
        fwd = [r["err"] for r in rs if "[" not in r["name"] or r["name"].startswith(("logits", "stat", "rollout"))]
        print(f"| {t} | {len(rs)} | {w['err']:.4f} | {w['bound']:.3g} | `{w['name']}` | "
              f"{max(norms) if norms else float('nan'):.4f} | {max(fwd) if fwd else float('nan'):.4f} |")
    print("\n## Gradient / moment tensors above 2e-2 (element-wise), all groups\n")
    print("| test group | tensor | error | bound |\n|---|---|---:|---:|")
    for r in sorted(recs, key=lambda r: -r["err"]):
        if "[elem" in r["name"] and r["err"] > 2e-2 and "noise floor" not in r["name"]:
            print(f"| {r['test']} | `{r['name']}` | {r['err']:.4f} | {r['bound']:.3g} |")
    nf = [r for r in recs if r["test"].startswith("noise floor")]
    if nf:
        print("\n## bf16 noise floor: the CUDA path next to stock `torch.autocast(bfloat16)` on the same tensors\n")
        print("| model | tensors | cuda path: max / mean element-wise error | stock autocast: max / mean | "
              "tensors where ours is noisier |\n|---|---:|---|---|---:|")
        for t in sorted({r["test"] for r in nf}):
            ours = [r["err"] for r in nf if r["test"] == t and "cuda path" in r["name"] and "[elem" in r["name"]]
            auto = [r["err"] for r in nf if r["test"] == t and "stock autocast" in r["name"] and "[elem" in r["name"]]
            frac = [r["err"] for r in nf if r["test"] == t and r["name"].startswith("fraction")]
            print(f"| {t} | {len(ours)} | {max(ours):.4f} / {sum(ours) / len(ours):.4f} | {max(auto):.4f} / "
                  f"{sum(auto) / len(auto):.4f} | {frac[0] * 100 if frac else float('nan'):.0f} % |")


if __name__ == "__main__":
    main()
