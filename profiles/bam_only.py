"""Runs only the BAM-records leg of bench.py (bam_path) -- for host-pipeline timing (BSGPU_TIMING=1) and profiling."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from bs_call_b200 import lib as bslib

ap = argparse.ArgumentParser()
ap.add_argument("--bam-sites", type=float, default=8e6)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--no-cpu", action="store_true")
args = ap.parse_args()
torch.cuda.set_device(0)
gpu = bslib.BsGpu(device=0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
out = bench.bam_path(args, gpu, bslib, torch, np, st.cuda_stream, 0, 1, 0)
print(json.dumps(out))
gpu.close()
