#!/bin/bash
# ncu evidence for bench.py (run under gpurun, 1 GPU). Outputs go to gpurun_out/.
set -u
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 0"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -c 300 gpurun_out/plain.log
# (1) every launch with its device time, skipping model build + first warm-up steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 1400 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
# (2) full sections for the dominant kernels: AdamW (HBM-bound), the tcgen05 pair GEMM, the XiT attention
ncu --set full --clock-control none --import-source on -k regex:adamw_multi_kernel -s 4 -c 2 \
    -o gpurun_out/prof_adamw -f $CMD > gpurun_out/ncu_adamw.log 2>&1
echo "adamw rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 40 -c 10 \
    -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:xattn_tc -s 8 -c 4 \
    -o gpurun_out/prof_xattn -f $CMD > gpurun_out/ncu_xattn.log 2>&1
echo "xattn rc=$?"
ls -la gpurun_out/*.ncu-rep
