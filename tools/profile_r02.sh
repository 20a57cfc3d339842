#!/bin/bash
# Round-2 ncu evidence (run under gpurun, 1 GPU).  Outputs go to gpurun_out/.
set -u
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-other-configs --profile-steps 0"
$CMD > gpurun_out/r2_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_plain.log; exit 1; }
tail -c 300 gpurun_out/r2_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 1400 --csv \
    --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
echo "launch list rc=$?"
python tools/prof_ndcg.py > gpurun_out/r2_prof_ndcg_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ndcg_warp32 -s 1 -c 1 \
    -o gpurun_out/r2_prof_ndcg -f python tools/prof_ndcg.py > gpurun_out/r2_ncu_ndcg.log 2>&1
echo "ndcg rc=$?"
ncu --set full --clock-control none --import-source on -k regex:adamw_multi_kernel -s 4 -c 2 \
    -o gpurun_out/r2_prof_adamw -f $CMD > gpurun_out/r2_ncu_adamw.log 2>&1
echo "adamw rc=$?"
ls -la gpurun_out/r2_*.ncu-rep
