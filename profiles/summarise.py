#!/usr/bin/env python
"""Turn the scratch outputs of profiles/run_profile.sh (gpurun_out/) into the tracked, per-round summaries here.
usage: python profiles/summarise.py r01b      (reads gpurun_out/<tag>_*; the first round's files had no tag prefix)"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


def launches():
    rows = [r for r in csv.reader(open(os.path.join(G, tag + "_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("bsgpu::", "").replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, r[8], r[7]])
        a[0] += 1
        a[1] += float(r[-1]) / 1e3
    tot = sum(a[1] for a in agg.values())
    lines = ["# kernel launches of `python bench.py --sites 3.2e7 --steps 1 --no-cpu --e2e-sites 1e6 --fused-sites 8e6 --bam-sites 2e6` under",
             "# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES, not absolutes)",
             "", "| kernel | launches | total us | share | grid | block |", "|---|---:|---:|---:|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| %s | %d | %.1f | %.1f%% | %s | %s |" % (k, a[0], a[1], 100 * a[1] / tot, a[2], a[3]))
    open(os.path.join(HERE, tag + "_launches.md"), "w").write("\n".join(lines) + "\n")
    shutil.copy(os.path.join(G, tag + "_launches.csv"), os.path.join(HERE, tag + "_launches.csv"))


def kernel(rep, name, per_site_n):
    txt = subprocess.run([sys.executable, os.path.join(HERE, "ncu_summary.py"), os.path.join(G, rep)], capture_output=True, text=True).stdout
    hot = subprocess.run([sys.executable, os.path.join(HERE, "ncu_source_hot.py"), os.path.join(G, rep), "25"], capture_output=True, text=True).stdout
    open(os.path.join(HERE, "%s_%s.txt" % (tag, name)), "w").write(
        "# ncu --set full --clock-control none --import-source on, read with profiles/ncu_summary.py and ncu_source_hot.py\n" + txt +
        "\n# hottest source lines (share of issued warp instructions, active threads per instruction, stall samples)\n" + hot)
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    def col(n):
        for i, h in enumerate(hdr):
            if h.endswith(n):
                return i
    rd, wr, tm = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
    def to_bytes(v, u):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    r = data[-1]
    b = to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
    return {"dram_bytes_per_launch": b, "sites_per_launch": per_site_n, "dram_bytes_per_site": b / per_site_n,
            "gpu_time_us": float(r[tm]) * ({"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}[units[tm]]),
            "note": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture; writes still resident in L2 at kernel end are not counted"}


launches()
traffic = {"k_call_sites": kernel(tag + "_prof_call.ncu-rep", "k_call_sites", 2_000_000),
           "k_pileup_tile": kernel(tag + "_prof_pile.ncu-rep", "k_pileup_tile", 8_000_000)}
if os.path.exists(os.path.join(G, tag + "_prof_reader.ncu-rep")):
    kernel(tag + "_prof_reader.ncu-rep", "reader_kernels", 1)
if os.path.exists(os.path.join(G, tag + "_prof_writer.ncu-rep")):
    kernel(tag + "_prof_writer.ncu-rep", "writer_kernels", 1)
# the headline kernel at bench size: dram bytes of a 62.5 M-site launch (ncu --metrics, one pass; run_profile_r02.sh)
tb = os.path.join(G, tag + "_traffic_call_bench.csv")
if os.path.exists(tb):
    rows = [r for r in csv.reader(open(tb)) if len(r) > 10 and r[0].isdigit()]
    per = collections.defaultdict(dict)
    for r in rows:
        name, unit, val = r[-3], r[-2], float(r[-1].replace(",", ""))
        per[r[0]][name] = val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1)
    launches_ = [v for v in per.values() if "dram__bytes_read.sum" in v]
    if launches_:
        b = sum(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for v in launches_) / len(launches_)
        traffic["k_call_sites_bench"] = {"dram_bytes_per_launch": b, "sites_per_launch": 62_500_000, "dram_bytes_per_site": b / 62_500_000,
                                         "gpu_time_us": sum(v.get("gpu__time_duration.sum", 0.0) for v in launches_) / len(launches_), "launches_averaged": len(launches_),
                                         "note": "dram__bytes_read.sum + dram__bytes_write.sum of bench-sized launches (python bench.py --sites 1e9), ncu --metrics, one pass"}
        shutil.copy(tb, os.path.join(HERE, tag + "_traffic_call_bench.csv"))
traffic["round"] = tag
json.dump(traffic, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
for f in ("bench_full.json", "bench_reference.json"):
    if os.path.exists(os.path.join(G, tag + "_" + f)):
        shutil.copy(os.path.join(G, tag + "_" + f), os.path.join(HERE, "%s_%s" % (tag, f)))
print(open(os.path.join(HERE, tag + "_launches.md")).read())
print(json.dumps(traffic, indent=1))
