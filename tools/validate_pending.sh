#!/bin/bash
# One-call GPU validation of everything DESIGN.md 6a lists as "written but not yet run on a GPU" (1 GPU part).
#   gpurun --timeout 900 -- 'bash tools/validate_pending.sh'
# Every check runs under its own timeout; the summary at the end says which ones passed.  2-GPU checks:
#   gpurun --gpus 2 --timeout 900 -- 'python -m pytest tests/test_dp_shard_gpu.py -q; python -m torch.distributed.run \
#     --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dp_stage2_check.py'
set -u
OUT=gpurun_out/pending
mkdir -p $OUT
declare -A RC
run() { name=$1; shift; timeout 300 "$@" > $OUT/$name.log 2>&1; RC[$name]=$?; tail -2 $OUT/$name.log; }

LR2_UNVALIDATED=1 run surrogate python -m pytest tests/test_surrogate_gpu.py -x -q
LR2_NDCG_BLOCK_PAIRS=1 LR2_NDCG_LEGACY=1 run ndcg_block_pairs python -m pytest tests/test_rows_gpu.py -k ndcg -x -q
LR2_WGRAD_ADAMW_IMPL=mma run wgrad_mma_unit python tools/check_wgrad_mma.py
run wgrad_tcgen05_unit python tools/check_wgrad_mma.py
LR2_WGRAD_ADAMW_IMPL=mma run wgrad_mma_stage3 python -m pytest tests/test_stage3_gpu.py -k fused_fc1 -x -q
LR2_WGRAD_ADAMW_IMPL=mma run bench_fused_mma python bench.py --fused-fc1 --steps 20 --warmup 5 --no-cpu-baseline --profile-steps 0
run bench_default python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-steps 0
echo "---- summary (0 = passed) ----"
for k in "${!RC[@]}"; do echo "$k rc=${RC[$k]}"; done
