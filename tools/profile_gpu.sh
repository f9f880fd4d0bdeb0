#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full ncu capture of the dominant kernel of bench.py.
# Usage: tools/profile_gpu.sh <tag> [kernel-regex]
TAG=${1:-r01}
KRE=${2:-phx_adj_kernel}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 60 -c 2 \
    -f -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 > gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
