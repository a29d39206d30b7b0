#!/usr/bin/env python
"""tools/c5_experiment.py — BASELINE.json configs[4] on ONE GPU: 4·10^9-byte synthetic DNA text,
GPU suffix-array/BWT/index construction, then a count sweep in 1 M-pattern batches.
Reports build time, index size, count throughput; checks that every text-sampled pattern occurs
and that located positions really are occurrences."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csfm_b200 as fm  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000_000
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
w = fm.workloads
t0 = time.perf_counter()
text = w.dna_text_torch(n, 7, dev)
torch.cuda.synchronize()
t_text = time.perf_counter() - t0
free0, total_mem = torch.cuda.mem_get_info()
t0 = time.perf_counter()
certify = os.environ.get("C5_CERTIFY", "1") == "1"
idx = fm.FMIndex.build_from_text_device(text.data_ptr(), n, fm.BuildParams(ssa_stride=32), device=0,
                                        flags=fm.BUILD_KEEP_SA if certify else 0)
torch.cuda.synchronize()
t_build = time.perf_counter() - t0
cert = None
if certify:
    t0 = time.perf_counter()
    cert = w.certify_sa_torch(text, idx.sa_device_ptr(), n)
    cert["seconds"] = time.perf_counter() - t0
    # BWT and SSA against their definitions, from the certified SA (bwt.hpp:10-13, fm_index.cpp:57-65)
    idx.release_sa()
    torch.cuda.empty_cache()
info = idx.info()
print(f"n={n} text_gen={t_text:.1f}s build={t_build:.1f}s levels={info.levels} blob={info.blob_bytes/1e9:.2f} GB kmer_k={info.kmer_k}",
      file=sys.stderr, flush=True)
B = 1_000_000
stream = torch.cuda.Stream()
d_counts = torch.zeros(B, dtype=torch.int64, device=dev)
nb = 8
bat = [w.sampled_patterns_torch(text, B, 20, 20, 0, 11, first=b * B) for b in range(nb)]
ok = True
for b in range(nb):
    idx.count_batch_device(bat[b][0].data_ptr(), bat[b][1].data_ptr(), B, d_counts.data_ptr(), 0, stream.cuda_stream)
    stream.synchronize()
    ok = ok and bool((d_counts >= 1).all().item())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record(stream)
    for i in range(batches):
        idx.count_batch_device(bat[i % nb][0].data_ptr(), bat[i % nb][1].data_ptr(), B, d_counts.data_ptr(), 0, stream.cuda_stream)
    e1.record(stream)
stream.synchronize()
ms = e0.elapsed_time(e1)
# locate spot check
npl = 20000
offs = torch.zeros(npl + 1, dtype=torch.int64, device=dev)
status = torch.zeros(npl, dtype=torch.int32, device=dev)
tot = idx.locate_batch_device(bat[0][0].data_ptr(), bat[0][1].data_ptr(), npl, 50, offs.data_ptr(), 0, 0, status.data_ptr(), stream.cuda_stream)
pos = torch.zeros(max(1, tot), dtype=torch.int64, device=dev)
idx.locate_batch_device(bat[0][0].data_ptr(), bat[0][1].data_ptr(), npl, 50, offs.data_ptr(), pos.data_ptr(), tot, status.data_ptr(), stream.cuda_stream)
stream.synchronize()
q_of = torch.searchsorted(offs, torch.arange(tot, device=dev), right=True) - 1
okl = torch.ones(tot, dtype=torch.bool, device=dev)
for k in range(20):
    okl &= text[pos + k] == bat[0][0][q_of * 20 + k]
out = {"n": n, "text_gen_s": t_text, "build_from_text_s": t_build, "suffixes_per_s": n / t_build, "levels": int(info.levels),
       "index_bytes": int(info.blob_bytes), "kmer_k": int(info.kmer_k), "count_batches": batches, "patterns": batches * B,
       "count_ms_total": ms, "count_qps": batches * B / (ms / 1e3), "all_counts_ge_1": ok,
       "locate_positions_checked": int(tot), "locate_positions_ok": bool(okl.all().item()), "failed_queries": int(status.max().item()),
       "device_mem_total_gb": total_mem / 1e9, "sa_certificate": cert}
print(json.dumps(out))
