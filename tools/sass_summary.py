"""`cuobjdump -sass` of the built library, reduced to the mnemonics that prove which hardware paths each kernel uses
(B200_PROFILING.md): UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = TMA
load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = mma.sync, VIMNMX = 32-bit min/max.
    python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lr2ppo_b200", "liblr2ppo_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "VIMNMX", "SHFL",
         "LDGSTS", "STG", "LDG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if cur and m:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["_total"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonic counts per kernel (`cuobjdump -sass lr2ppo_b200/liblr2ppo_b200.so`, sm_100a)\n")
    print("Static instruction counts of the shipped binary; a column is empty when the kernel has none.\n")
    print("| kernel | instr | " + " | ".join(WATCH) + " |")
    print("|---|---:|" + "---:|" * len(WATCH))
    tot = collections.Counter()
    for (name, c), nice in zip(kernels.items(), demangle):
        nice = re.sub(r"\(.*", "", nice).replace("lr2::", "")
        print(f"| `{nice}` | {c['_total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WATCH) + " |")
        tot.update(c)
    print(f"| **all {len(kernels)} kernels** | {tot['_total']} | " + " | ".join(str(tot[w]) if tot[w] else "" for w in WATCH) + " |")


if __name__ == "__main__":
    main()
