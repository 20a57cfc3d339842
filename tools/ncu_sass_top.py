"""Top stall sites of one kernel from an `ncu --set full --import-source on` report, per SASS instruction.
    python tools/ncu_sass_top.py gpurun_out/r2_prof_gemm_plain.ncu-rep [rows=30] > profiles/r02_gemm_plain_sass.md"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall = [i for i, nm in enumerate(h) if nm.startswith("stall_") and "Not Issued" not in nm]
    body = [r for r in rows[2:] if len(r) > isamp and r[isamp].isdigit()]
    tot = sum(int(r[isamp]) for r in body)
    agg = {}
    for r in body:
        for i in stall:
            if r[i].isdigit():
                agg[h[i][6:]] = agg.get(h[i][6:], 0) + int(r[i])
    print(f"# {rep}\n\n`{rows[0][1][:140]}`\n")
    print(f"{tot} warp-state samples, {sum(int(r[iex]) for r in body)} warp instructions executed.\n")
    print("Stall reasons over all samples: " + ", ".join(f"{k} {v} ({100 * v / tot:.0f} %)" for k, v in
                                                        sorted(agg.items(), key=lambda kv: -kv[1])[:8]) + "\n")
    print("| SASS # | samples | share | executed | instruction | top stall reasons |\n|---:|---:|---:|---:|---|---|")
    for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][isamp]))[:n]:
        st = {h[i][6:]: int(r[i]) for i in stall if r[i].isdigit() and int(r[i]) > 0}
        st = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"| {idx} | {r[isamp]} | {100 * int(r[isamp]) / tot:.1f} % | {r[iex]} | `{r[ia].strip()[:70]}` | {st} |")


if __name__ == "__main__":
    main()
