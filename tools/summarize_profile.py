"""Turn gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep into the small text summaries committed under
profiles/.   Usage: python tools/summarize_profile.py <tag>"""
import collections
import csv
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(REPO, "profiles")
os.makedirs(out_dir, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__cycles_active.avg", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]

lp = os.path.join(REPO, "gpurun_out", "launches_%s.csv" % tag)
if os.path.exists(lp):
    lines = [l for l in open(lp) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else (v * 1e6 if u == "s" else v))
        a = agg[row["Kernel Name"][:70]]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, "%s_launches.txt" % tag), "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none, %s\n" %
                 (sys.argv[2] if len(sys.argv) > 2 else "python bench.py --steps 2 --warmup 3"))
        fh.write("# per-launch times are cold-cache and serialised: compare SHARES\n")
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            fh.write("%-72s n=%4d total_us=%11.1f avg_us=%9.1f share=%5.1f%%\n" % (n, c, t, t / c, 100 * t / tot))
    print(open(os.path.join(out_dir, "%s_launches.txt" % tag)).read())

import glob
for rp in sorted(glob.glob(os.path.join(REPO, "gpurun_out", "prof_%s*.ncu-rep" % tag))):
    sub = os.path.basename(rp)[len("prof_"):-len(".ncu-rep")]
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = {}
    with open(os.path.join(out_dir, "%s_ncu_full.txt" % sub), "w") as fh:
        fh.write("# ncu --set full --clock-control none --import-source on (selected raw metrics per captured launch)\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            fh.write("== %s\n" % name)
            vals = {}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    fh.write("   %-70s %s %s\n" % (k, r[i], units[i]))
                    vals[k] = (r[i], units[i])
            def to_bytes(k):
                v, u = vals[k]
                v = float(v.replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            if "dram__bytes_read.sum" in vals:
                key = name.split("(")[0].split("::")[-1].split("<")[0].strip()
                traffic.setdefault(key, []).append(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
    print(open(os.path.join(out_dir, "%s_ncu_full.txt" % sub)).read())
    tj = os.path.join(out_dir, "traffic.json")
    cur = json.load(open(tj)) if os.path.exists(tj) else {}
    for k, v in traffic.items():
        cur[k] = sum(v) / len(v)
    cur["_source"] = "profiles/*_ncu_full.txt (dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    json.dump(cur, open(tj, "w"), indent=1)
