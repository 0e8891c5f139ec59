"""Ceiling of the rebuilding pass alone (no GPU activity): bsgpu_wire_expand over the pool size, on the box's cores."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bs_call_b200 import lib
L = lib.load()
n = 16_000_000
w = np.random.default_rng(0).integers(0, 255, (n, 120), dtype=np.uint8)
out = np.zeros(n * 200, dtype=np.uint8); sk = np.zeros(n, dtype=np.uint8)
for T in (1, 4, 8, 12, 16):
    best = 1e9
    for rep in range(3):
        t = time.time()
        L.bsgpu_wire_expand(w.ctypes.data_as(C.c_void_p), C.c_size_t(n), C.c_size_t(200), out.ctypes.data_as(C.c_void_p), sk.ctypes.data_as(C.c_void_p), C.c_int(T))
        best = min(best, time.time() - t)
    print(T, "threads %.1f M sites/s %.1f GB/s" % (n / best / 1e6, n * 320 / best / 1e9), flush=True)
# plain numpy copy bandwidth on one thread for scale
a = np.zeros(1 << 30, dtype=np.uint8); b = np.zeros(1 << 30, dtype=np.uint8)
t = time.time(); b[:] = a; d = time.time() - t
print("memcpy 1 thread %.1f GB/s (read+write)" % (2 * (1 << 30) / d / 1e9))
