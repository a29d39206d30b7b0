"""GPU, BASELINE.json configs[1] at FULL size (C2: n = 2^26 DNA + '$', 1 M text-sampled patterns of
length 20): the GPU-built suffix array is certified by the oracle's O(n) checker (a certified SA is
THE reference SA: suffixes are distinct and the order is total), BWT / SSA are recomputed from it on
the host, and all 1 M counts and intervals are compared with the oracle over that SA. Locate is
checked through its defining property on every reported position."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    import csfm_b200 as fm
    w = fm.workloads
    n = 1 << 26
    text = w.dna_text_np(n, 1)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=32), flags=fm.BUILD_KEEP_SA)
    sa = idx.sa()
    idx.release_sa()
    return fm, text, idx, sa


def test_c2_build_products(c2):
    fm, text, idx, sa = c2
    n = text.size
    assert oracle.sa_check(text, sa) == 0                      # certificate
    bwt = text[(sa.astype(np.int64) - 1) % n]                  # bwt.hpp:10-13
    assert (idx.bwt() == bwt).all()
    assert (idx.ssa() == sa[::32]).all()                       # fm_index.cpp:57-65
    C = np.zeros(257, np.uint32)
    C[1:] = np.cumsum(np.bincount(text, minlength=256)).astype(np.uint32)
    assert (idx.C_array() == C).all()
    info = idx.info()
    assert info.levels == 1 and info.layout == 3 and info.line_bytes == 64 and info.sigma == 5   # two-bit symbols: csfm_dna.cuh


def test_c2_one_million_counts_vs_oracle(c2):
    fm, text, idx, sa = c2
    d, o = fm.workloads.sampled_patterns_np(text, 1_000_000, 20, 20, 0, 2)
    counts, sp_ep = idx.count_batch(d, o, want_intervals=True)
    orc = oracle.OracleIndex(text, stride=32, sa=sa)
    oc, ose = orc.count_batch(d, o)
    assert (counts == oc).all()
    assert (sp_ep == ose).all()
    assert (counts >= 1).all()
    # secondary set: uniform random strings (mostly misses, early exits)
    d2, o2 = fm.workloads.random_patterns_np(b"ACGT", 200_000, 20, 9)
    c2_, se2 = idx.count_batch(d2, o2, want_intervals=True)
    oc2, ose2 = orc.count_batch(d2, o2)
    assert (c2_ == oc2).all() and (se2 == ose2).all()
    # locate on a slice of the batch: identical to the oracle, and every position is an occurrence
    sub = 20_000
    offs, pos, status = idx.locate_batch(d[: 20 * sub], o[: sub + 1], limit=100)
    ooffs, opos, ostatus, _ = orc.locate_batch(d[: 20 * sub], o[: sub + 1], limit=100)
    assert (offs == ooffs).all() and (status == ostatus).all() and (pos == opos).all()
    q_of = np.repeat(np.arange(sub), np.diff(offs).astype(np.int64))
    for k in range(20):
        assert (text[pos.astype(np.int64) + k] == d[q_of * 20 + k]).all()
