#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
L=gpurun_out/tc_exp.log
: > $L
echo "== prof" >> $L
PHX_TC_PROF=1 timeout 300 python tools/tc_check.py --shapes 20000,200,4096 --reps 3 --modes 3xtf32 >> $L 2>&1
echo "== plain" >> $L
timeout 300 python tools/tc_check.py --shapes 350,40,10000 3551,120,1024 11165,200,1024 20000,200,4096 20000,256,4096 --reps 5 --modes 3xtf32 tf32 >> $L 2>&1
echo "== ncu launch list (20000,200,4096)" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tc_launches.csv python tools/tc_check.py --shapes 20000,200,4096 --reps 2 --modes 3xtf32 tf32 > /dev/null 2>&1
grep -E "tc_|sgemm" gpurun_out/tc_launches.csv | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | tail -6 >> $L
echo "== pytest" >> $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $L
cat $L
