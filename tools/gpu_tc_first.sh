#!/bin/bash
# first GPU contact of the tcgen05 batched RHS: correctness in both descriptor variants, then timings, then the suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/tc_first.log 2>&1
echo "== DEBUG=0 small" >> gpurun_out/tc_first.log
PHX_TC_DEBUG=0 timeout 180 python tools/tc_check.py --shapes 350,40,256 1001,100,300 --reps 3 >> gpurun_out/tc_first.log 2>&1
echo "rc=$?" >> gpurun_out/tc_first.log
echo "== DEBUG=1 small" >> gpurun_out/tc_first.log
PHX_TC_DEBUG=1 timeout 180 python tools/tc_check.py --shapes 350,40,256 --reps 3 --modes 3xtf32 >> gpurun_out/tc_first.log 2>&1
echo "rc=$?" >> gpurun_out/tc_first.log
echo "== DEBUG=0 large" >> gpurun_out/tc_first.log
PHX_TC_DEBUG=0 timeout 300 python tools/tc_check.py --shapes 3551,120,1024 11165,200,1024 20000,200,4096 --reps 5 >> gpurun_out/tc_first.log 2>&1
echo "rc=$?" >> gpurun_out/tc_first.log
echo "== pytest gpu" >> gpurun_out/tc_first.log
timeout 900 python -m pytest tests -m gpu -x -q >> gpurun_out/tc_first.log 2>&1
echo "rc=$?" >> gpurun_out/tc_first.log
tail -40 gpurun_out/tc_first.log
