#!/usr/bin/env python
"""The deep-panel pileup (500x, config 4 shape) on its own, for ncu: python profiles/deep_case.py [sites] [depth]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bs_call_b200 import lib  # noqa: E402

sz = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
depth = float(sys.argv[2]) if len(sys.argv) > 2 else 500.0
gpu = lib.BsGpu()
L = 150
ns = gpu.synth_block_nseg(sz, L, depth)
seg = torch.empty(ns * 16 + 16, dtype=torch.uint8, device="cuda")
b = torch.empty(ns * L + 16, dtype=torch.uint8, device="cuda")
r = torch.empty(sz + 16, dtype=torch.uint8, device="cuda")
p = torch.empty(sz * 104 + 16, dtype=torch.uint8, device="cuda")
gpu.synth_block_dev(20261018, 1000, sz, L, depth, seg.data_ptr(), ns, b.data_ptr(), ns * L, r.data_ptr())
gpu.sync()
for _ in range(3):
    gpu.pileup_block_dev(seg.data_ptr(), ns, b.data_ptr(), 1000, sz, p.data_ptr())
gpu.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
gpu.pileup_block_dev(seg.data_ptr(), ns, b.data_ptr(), 1000, sz, p.data_ptr())
gpu.sync()
e1.record()
torch.cuda.synchronize()
print("sites %d depth %g segments %d" % (sz, depth, ns))
