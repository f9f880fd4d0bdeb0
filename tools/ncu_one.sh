#!/bin/bash
# One full ncu capture (with source correlation) of a resident kernel on the headline shape.
# Usage: tools/ncu_one.sh <tag> <kernel-regex> [G H method dt]
TAG=$1; KRE=$2; shift 2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 4 -c 1 -f -o gpurun_out/src_${TAG} \
    python tools/phase_profile.py "$@" > gpurun_out/ncu_one_${TAG}.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_one_${TAG}.log
