"""Bucket the warp-stall samples of an ncu source page by function (line ranges found by scanning the source for
top-level definitions).  Usage: python tools/ncu_buckets.py <report.ncu-rep> [source-file]"""
import csv
import re
import subprocess
import sys
import os

rep = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "phoenix_b200", "csrc", "phx_resident.cuh")
lines = open(src).read().split("\n")
starts = []
for i, l in enumerate(lines, 1):
    m = re.match(r"^(?:__device__|__global__|template|static|inline).*?\b([A-Za-z_0-9]+)\s*\(", l)
    if l.startswith("__device__") or l.startswith("__global__"):
        m = re.search(r"\b([A-Za-z_0-9]+)\s*\(", l.replace("__launch_bounds__(PHX_THREADS, 1)", ""))
        if m:
            starts.append((i, m.group(1)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None
fname = ""
agg = {}
tot = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci = {}
        for k, h in enumerate(hdr):
            ci.setdefault(h, k)
        continue
    if hdr is None or len(r) != len(hdr) or not r[0].isdigit():
        continue
    try:
        smp = int(r[ci["# Samples"]])
        inst = int(r[ci["Instructions Executed"]])
    except ValueError:
        continue
    ln = int(r[0])
    name = fname
    if fname == os.path.basename(src):
        name = "?"
        for st, fn in starts:
            if st <= ln:
                name = fn
    a = agg.setdefault(name, [0, 0, {}])
    a[0] += smp
    a[1] += inst
    for k, h in enumerate(hdr):
        if h.startswith("stall_") and "(" not in h and r[k].isdigit():
            a[2][h[6:]] = a[2].get(h[6:], 0) + int(r[k])
    tot += smp
print("total samples", tot)
for name, (smp, inst, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:30]:
    top = sorted(st.items(), key=lambda x: -x[1])[:3]
    print("%-28s %6d %5.1f%%  inst %9d   %s" % (name, smp, 100.0 * smp / tot, inst, " ".join("%s=%d" % kv for kv in top)))
