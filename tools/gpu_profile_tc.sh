#!/bin/bash
# Round profile of the tensor-core batched RHS: launch list + full captures of the two MMA kernels (BASELINE config 5 shape)
TAG=${1:-r01n}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
CMD="python tools/tc_check.py --shapes 20000,200,4096 --reps 3 --modes 3xtf32"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}_tc.csv $CMD > gpurun_out/tc_under_ncu_${TAG}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tc_branch -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_tc_branch $CMD >> gpurun_out/tc_under_ncu_${TAG}.log 2>&1
echo "branch capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tc_joint -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_tc_joint $CMD >> gpurun_out/tc_under_ncu_${TAG}.log 2>&1
echo "joint capture rc=$?"
