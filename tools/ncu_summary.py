"""Summarise ncu outputs into small, committed text files under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/ncu_summary.py rep gpurun_out/prof_gemm.ncu-rep profiles/r01_gemm_full.md
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("lr2::", "")


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="ignore")) if len(r) > 14 and r[0].isdigit()]
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = short(r[4])
        tot[k][0] += 1
        tot[k][1] += float(r[14]) / 1e3
    total = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({len(rows)} launches, gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {us:.1f} | {us / total * 100:.1f}% |\n")
        f.write(f"\ntotal {total:.1f} us\n")
    print(open(dst).read()[:3000])


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__cycles_active.avg", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def rep(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full capture: {src}\n\n")
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            f.write(f"## {name}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n\n")
            for m in hdr:
                if m in METRICS or m in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                                         "sm__inst_executed_pipe_tensor.sum",
                                         "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
                                         "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                                         "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
                                         "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
                                         "smsp__inst_executed.sum", "sm__cycles_elapsed.max"):
                    f.write(f"- {m} = {r[hdr.index(m)]} {units[hdr.index(m)]}\n")
            f.write("\n")
    print(open(dst).read()[:4000])


if __name__ == "__main__":
    {"launches": launches, "rep": rep}[sys.argv[1]](sys.argv[2], sys.argv[3])
