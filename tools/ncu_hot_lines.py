"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line: where the warp-stall
samples of a kernel are.   Usage: python tools/ncu_hot_lines.py <report.ncu-rep> [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
out = []
hdr = None
fname = ""
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
    stalls = {h[6:]: int(r[k]) for k, h in enumerate(hdr) if h.startswith("stall_") and "(" not in h and r[k].isdigit()}
    out.append((smp, inst, fname, r[0], r[1].strip()[:100], stalls))
tot = sum(o[0] for o in out) or 1
toti = sum(o[1] for o in out) or 1
print("total samples %d, warp instructions %d" % (tot, toti))
agg = {}
for o in out:
    for k, v in o[5].items():
        agg[k] = agg.get(k, 0) + v
print("stall mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
key = (lambda x: -x[1]) if len(sys.argv) > 3 else (lambda x: -x[0])
for smp, inst, f, ln, src, st in sorted(out, key=key)[:top]:
    main = sorted(st.items(), key=lambda x: -x[1])[:2]
    print("%6d %5.1f%% inst %5.1f%% %s:%s  %-100s %s" % (smp, 100.0 * smp / tot, 100.0 * inst / toti, f, ln, src,
                                                       " ".join("%s=%d" % kv for kv in main)))
