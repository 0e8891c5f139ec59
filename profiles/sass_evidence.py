#!/usr/bin/env python
"""SASS mnemonic counts per kernel of bs_call_b200/libbsgpu.so (cuobjdump -sass): python profiles/sass_evidence.py > profiles/<tag>_sass_evidence.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "bs_call_b200", "libbsgpu.so")], capture_output=True, text=True).stdout
kern, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").split("(")[0]
        cur = kern.setdefault(name.replace("bsgpu::", "").replace("void ", ""), collections.Counter())
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur["lines"] += 1
        cur[m.group(1).split(".")[0]] += 1
cols = [("UBLKCP", ["UBLKCP"]), ("SYNCS", ["SYNCS"]), ("ATOMS", ["ATOMS"]), ("ATOMG/RED", ["ATOMG", "RED"]), ("MATCH", ["MATCH"]), ("REDUX", ["REDUX"]),
        ("DFMA+DADD+DMUL", ["DFMA", "DADD", "DMUL"]), ("MUFU", ["MUFU"]), ("LDS", ["LDS"]), ("STS", ["STS"]), ("LDG", ["LDG"]), ("STG", ["STG"]), ("SHF", ["SHF"])]
print("# SASS evidence (cuobjdump -sass bs_call_b200/libbsgpu.so, sm_100a): instruction mnemonics per kernel.")
print("# UBLKCP = cp.async.bulk (TMA bulk copy engine), SYNCS = mbarrier arrive / try_wait, ATOMS = shared-memory atomics,")
print("# MATCH = match.any (warp-aggregated binning), REDUX = warp reduce, DFMA/DADD/DMUL = FP64 pipe, MUFU = special function unit,")
print("# SHF = funnel shift (the writer's word-stitching copy-out)\n")
print("| kernel | SASS lines | " + " | ".join(c for c, _ in cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
for k, c in kern.items():
    print("| %s | %d | %s |" % (k, c["lines"], " | ".join(str(sum(c[m] for m in ms)) for _, ms in cols)))
