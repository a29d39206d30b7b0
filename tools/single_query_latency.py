#!/usr/bin/env python3
"""Latency of ONE count() call through the C ABI (csfm_count_batch with npat == 1) on the C1 text: median / p95 in µs
over the 10 000 random length-5 patterns of tools/benchmark.cpp. CSFM_LIB selects an experiment build."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import csfm_b200 as fm  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "c1_workload.npz"))
text = z["text"]
idx = fm.FMIndex.build_from_text(text, fm.BuildParams())
L = fm.lib()
pats = [np.ascontiguousarray(text[p:p + 5]) for p in z["rand_pos"]]
offs = np.array([0, 5], np.uint64)
out = np.zeros(1, np.uint64)
fn = L.csfm_count_batch
h, po, pout = idx._h, C.c_void_p(offs.ctypes.data), C.c_void_p(out.ctypes.data)
ptrs = [C.c_void_p(p.ctypes.data) for p in pats]
for k in range(2000):
    fn(h, ptrs[k], po, 1, pout, None)
lat = np.empty(len(pats))
tot = 0
for k, p in enumerate(ptrs):
    t0 = time.perf_counter_ns()
    fn(h, p, po, 1, pout, None)
    lat[k] = time.perf_counter_ns() - t0
    tot += int(out[0])
print(f"{os.environ.get('CSFM_LIB', 'libcsfm.so')}: p50 {np.median(lat) / 1e3:.2f} us  p95 {np.percentile(lat, 95) / 1e3:.2f} us  "
      f"mean {lat.mean() / 1e3:.2f} us  total_matches {tot} (reference 9907582)")
