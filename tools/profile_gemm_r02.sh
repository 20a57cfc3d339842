#!/bin/bash
# Round-2 GEMM evidence (run under gpurun, 1 GPU): per-shape table, then source-level ncu captures of the pair kernel
# on the plain and the GELU+dropout forward shapes.  Outputs go to gpurun_out/.
set -u
python tools/gemm_bench.py > gpurun_out/r2_gemm_shapes_final.txt 2>&1 || { tail -5 gpurun_out/r2_gemm_shapes_final.txt; exit 1; }
cat gpurun_out/r2_gemm_shapes_final.txt
for d in 1 2 3; do LR2_GEMM_DBG=$d python tools/gemm_bench.py plain 2>&1 | grep plain | sed "s/^/DBG=$d /"; done | tee gpurun_out/r2_gemm_dbg_final.txt
ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 8 -c 1 \
    -o gpurun_out/r2_prof_gemm_plain -f python tools/gemm_bench.py plain > gpurun_out/r2_ncu_gemm_plain.log 2>&1
echo "plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 8 -c 1 \
    -o gpurun_out/r2_prof_gemm_geludrop -f python tools/gemm_bench.py "gelu+drop" > gpurun_out/r2_ncu_gemm_geludrop.log 2>&1
echo "gelu+drop rc=$?"
ls -la gpurun_out/r2_prof_gemm_*.ncu-rep
