#!/usr/bin/env python
"""tools/sanitizer_check.py — small end-to-end run for `compute-sanitizer --tool memcheck`:
builds two indexes (both layouts, one and two levels), runs count (direct and TMA-staged),
locate and BWT extraction, checks the answers against the oracle, prints OK."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csfm_b200 as fm  # noqa: E402
import oracle  # noqa: E402

rng = np.random.default_rng(5)
for sigma, n, flags, staging in [(4, 30_000, 0, "direct"), (200, 40_000, 0, "tma"), (30, 20_000, fm.BUILD_LAYOUT_BINARY64, "direct")]:
    os.environ["CSFM_PATTERN_STAGING"] = staging
    alpha = np.sort(rng.choice(np.arange(1, 256), sigma, replace=False)).astype(np.uint8)
    text = np.concatenate([alpha[rng.integers(0, sigma, n)], np.zeros(1, np.uint8)]).astype(np.uint8)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=8), flags=flags)
    orc = oracle.OracleIndex(text, stride=8)
    pats = [text[s:s + m].tobytes() for s, m in zip(rng.integers(0, n - 80, 3000), rng.integers(0, 70, 3000))]
    d, o = fm.pack_patterns(pats)
    c, se = idx.count_batch(d, o, want_intervals=True)
    oc, ose = orc.count_batch(d, o)
    assert (c == oc).all() and (se == ose).all()
    offs, pos, status = idx.locate_batch(d, o, limit=9)
    ooffs, opos, ostatus, _ = orc.locate_batch(d, o, limit=9)
    assert (offs == ooffs).all() and (pos == opos).all() and (status == ostatus).all()
    assert (idx.bwt() == orc.bwt).all()
    idx.close()
print("sanitizer_check OK")
