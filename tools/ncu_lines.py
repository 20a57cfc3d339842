"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: python tools/ncu_lines.py dump.csv [kernel_index=1] [min_pct=0.5]"""
import csv
import sys


def main():
    path = sys.argv[1]
    want = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    rows = list(csv.reader(open(path)))
    kern, cur, hdr, agg, src, name = 0, None, None, {}, {}, ""
    keys = {"s": "# Samples", "i": "Instructions Executed", "lsb": "stall_long_sb", "wait": "stall_wait",
            "noinst": "stall_no_inst", "math": "stall_math", "ssb": "stall_short_sb", "br": "stall_branch_resolving",
            "nsel": "stall_not_selected"}

    def num(v):
        try:
            return int(v)
        except ValueError:
            return 0
    for r in rows:
        if r and r[0] == "Function Name":
            kern += 1
            if kern == want:
                name = r[1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            col = {k: hdr.index(v) for k, v in keys.items()}
            continue
        if kern != want or hdr is None or len(r) < len(hdr) - 2:
            continue
        if r[0] != "":
            cur = int(r[0]); src[cur] = r[1]
            continue
        a = agg.setdefault(cur, dict.fromkeys(list(keys) + ["n"], 0))
        for k, c in col.items():
            a[k] += num(r[c])
        a["n"] += 1
    tot = sum(a["s"] for a in agg.values()) or 1
    ti = sum(a["i"] for a in agg.values()) or 1
    print(name[:100])
    print("total samples", tot, "warp instructions", ti)
    print("line  samples%  inst%  sass  long_sb wait no_inst math short_sb branch not_sel | source")
    for ln in sorted(agg):
        a = agg[ln]
        if a["s"] * 100 >= tot * min_pct:
            print(f"{ln:4d} {100 * a['s'] / tot:6.1f}% {100 * a['i'] / ti:6.1f}% {a['n']:5d} {a['lsb']:6d} {a['wait']:5d} "
                  f"{a['noinst']:6d} {a['math']:5d} {a['ssb']:6d} {a['br']:5d} {a['nsel']:6d} | {src.get(ln, '')[:90]}")


if __name__ == "__main__":
    main()
