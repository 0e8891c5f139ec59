"""The host-buffer leg of the headline metric alone (bsgpu_call_sites on pinned arrays): for tuning the chunk size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bs_call_b200 import lib as bslib
from bs_call_b200.records import GT_METH, PILEUP
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 8000000
gpu = bslib.BsGpu(device=0)
d_p = torch.empty(n * 104 + 16, dtype=torch.uint8, device="cuda")
d_r = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
gpu.synth_sites_dev(20251018, 0, n, 30.0, d_p.data_ptr(), d_r.data_ptr(), 0)
torch.cuda.synchronize()
hp, hr, ho, hs = bslib.HostBuffer(n, PILEUP), bslib.HostBuffer(n, np.uint8), bslib.HostBuffer(n, GT_METH), bslib.HostBuffer(n, np.uint8)
hp.array.view(np.uint8)[:] = d_p[:n * 104].cpu().numpy(); hr.array[:] = d_r[:n].cpu().numpy()
for _ in range(2):
    gpu.call_sites(hp.array, hr.array, out=ho.array, skip=hs.array)
t0 = time.perf_counter()
for _ in range(5):
    gpu.call_sites(hp.array, hr.array, out=ho.array, skip=hs.array)
dt = (time.perf_counter() - t0) / 5
import zlib
st = gpu.stats()
print("chunk", os.environ.get("BSGPU_SITES_CHUNK", "default"), "wire", os.environ.get("BSGPU_WIRE", "default"), "threads", os.environ.get("BSGPU_EXPAND_THREADS", "default"),
      "cores", os.cpu_count(), "sites/s %.1f M" % (int((hp.array["n"] > 0).sum()) / dt / 1e6), "ms %.2f" % (dt * 1e3),
      "wire share %.2f" % (st["wire_sites"] / max(1, st["sites"])), "crc", zlib.crc32(ho.array.view(np.uint8)), zlib.crc32(hs.array))
if os.environ.get("E2E_ONLY_SITES"):
    sys.exit(0)
refw = np.concatenate([hr.array, [1, 1]]).astype(np.uint8)
hrw = bslib.HostBuffer(n + 2, np.uint8); hrw.array[:] = refw
hb = bslib.HostBuffer(n * 160, np.uint8)
for _ in range(2):
    gpu.call_sites_bcf(hp.array, hrw.array, 1, out=hb.array)
t0 = time.perf_counter()
for _ in range(5):
    b, r = gpu.call_sites_bcf(hp.array, hrw.array, 1, out=hb.array)
dt = (time.perf_counter() - t0) / 5
print("  to BCF records: sites/s %.1f M" % (int((hp.array["n"] > 0).sum()) / dt / 1e6), "ms %.2f" % (dt * 1e3), "H2D GB/s %.1f" % (n * 105 / dt / 1e9), "records", r)
