#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python tools/tc_check.py --vjp --shapes 350,40,10000 3551,120,1024 11165,200,1024 20000,200,4096 --reps 3 --modes 3xtf32 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vjp_launches.csv python tools/tc_check.py --vjp --shapes 20000,200,4096 --reps 1 --modes 3xtf32 > /dev/null 2>&1
python - <<PY
import csv
rows=[l for l in open("gpurun_out/vjp_launches.csv") if not l.startswith("==")]
r=list(csv.DictReader(rows))
for x in r[-16:]:
    print(x["Kernel Name"][:60], x["Metric Value"])
PY
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
