#!/usr/bin/env python
"""Aggregate an ncu source page (ncu -i X --page source --csv --print-source cuda,sass) by CUDA source line:
instructions executed, samples, dominant stall reasons.  usage: ncu_source_hot.py report.ncu-rep [kernel-index] [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
SORTKEY = sys.argv[3] if len(sys.argv) > 3 else "samples"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# the dump is a sequence of per-file sections: ("File Path", p), ("Function Name", f), header row, data rows
agg = collections.OrderedDict()
tot_inst = 0
hdr = None
fpath = None
seen_first_kernel = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name":
        fn = r[1]
        if seen_first_kernel is None: seen_first_kernel = fn
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    try:
        ln = int(d["Line No"])
    except Exception:
        continue
    key = (fpath.split("/")[-1], ln)
    a = agg.setdefault(key, dict(src=r[1], inst=0, tinst=0, samples=0, stalls=collections.Counter()))
    def num(x):
        try: return float(x)
        except Exception: return 0.0
    a["inst"] += num(d.get("Instructions Executed", 0)); a["tinst"] += num(d.get("Thread Instructions Executed", 0)); a["samples"] += num(d.get("# Samples", 0))
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k: a["stalls"][k] += num(d[k])
# rows repeat per kernel instance of the same function; totals are still proportional
tot_inst = sum(a["inst"] for a in agg.values()); tot_s = sum(a["samples"] for a in agg.values())
print("total inst %.3g samples %d" % (tot_inst, tot_s))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][SORTKEY])[:top]:
    st = ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(a["samples"], 1)) for k, v in a["stalls"].most_common(3))
    print("%-22s:%-4d inst %5.1f%% thr/inst %4.1f samples %5.1f%%  [%s]  %s" % (key[0], key[1], 100 * a["inst"] / max(tot_inst, 1), a["tinst"] / max(a["inst"], 1), 100 * a["samples"] / max(tot_s, 1), st, a["src"].strip()[:70]))
