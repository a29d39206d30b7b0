"""GPU: the suffix-array builder (csrc/csfm_sa.cu: packed-key radix sort, then prefix doubling that re-sorts only
the suffixes whose group still has more than one member) against the reference's order
(build_sa_naive, /root/reference/src/core/sais.hpp:8-16: unsigned bytes, a proper prefix first).

Small and medium texts are compared with the oracle's suffix array (itself pinned to build_sa_naive in
tests/test_oracle_vs_reference.py); the 2^28-byte text with planted 10^5-byte repeats is certified by the O(n)
checker (a certified SA is THE reference SA)."""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fm():
    import csfm_b200
    csfm_b200.lib()
    return csfm_b200


def _texts():
    rng = np.random.default_rng(5)
    out = {}
    out["all_same"] = np.full(40_000, 0x61, np.uint8)
    out["all_same_term"] = np.concatenate([np.full(40_000, 0x61, np.uint8), np.zeros(1, np.uint8)])
    out["period2"] = np.tile(np.frombuffer(b"ab", np.uint8), 30_000)
    out["period7_term"] = np.concatenate([np.tile(np.frombuffer(b"abcabca", np.uint8), 9_000), np.zeros(1, np.uint8)])
    body = rng.integers(1, 5, 300_000).astype(np.uint8)
    block = body[1000:6000].copy()
    for at in (50_000, 120_000, 200_001, 290_000):     # planted 5000-byte repeats, one of them cut by the text end
        m = min(block.size, body.size - at)
        body[at:at + m] = block[:m]
    out["planted_repeats_term"] = np.concatenate([body, np.zeros(1, np.uint8)])
    fib = [b"a", b"ab"]
    while len(fib[-1]) < 200_000:
        fib.append(fib[-1] + fib[-2])
    out["fibonacci"] = np.frombuffer(fib[-1], np.uint8).copy()
    out["bytes_repeat"] = np.tile(rng.integers(0, 256, 777).astype(np.uint8), 150)
    return out


TEXTS = _texts()


@pytest.mark.parametrize("name", sorted(TEXTS))
def test_sa_equals_oracle_on_repetitive_texts(fm, name):
    text = TEXTS[name]
    want = oracle.sa_build(text)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=16), flags=fm.BUILD_KEEP_SA)
    info = idx.info()
    assert (idx.sa() == want).all()
    assert info.sa_rounds >= 1 and info.sa_pair_passes >= text.size
    # the dense-only builder (every round sorts all n pairs, the round-1 algorithm) must give the same array
    os.environ["CSFM_SA_DENSE_ONLY"] = "1"
    try:
        dense = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=16), flags=fm.BUILD_KEEP_SA)
    finally:
        del os.environ["CSFM_SA_DENSE_ONLY"]
    assert (dense.sa() == want).all()
    if info.sa_rounds > 2:
        assert info.sa_pair_passes < dense.info().sa_pair_passes   # the unresolved suffixes only


def test_sa_2_28_text_with_planted_long_repeats(fm):
    """n = 2^28 DNA text with three planted copies of a 10^5-byte block (one overlapping itself): the doubling needs
    log2(1e5 / 21) ~ 13 rounds, but from round 1 on only the ~3e5 suffixes inside the copies are sorted."""
    import time
    import torch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n = 1 << 28
    text = fm.workloads.dna_text_torch(n, 21, dev)
    block = text[1_000_000:1_100_000].clone()
    for at in (50_000_000, 150_000_003, 150_050_003):
        text[at:at + block.numel()] = block
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx = fm.FMIndex.build_from_text_device(text.data_ptr(), n, fm.BuildParams(ssa_stride=32), device=0, flags=fm.BUILD_KEEP_SA)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    info = idx.info()
    cert = fm.workloads.certify_sa_torch(text, idx.sa_device_ptr(), n)
    assert cert["ok"], cert
    print(f"\\nplanted repeats: n=2^28 rounds={info.sa_rounds} dense_passes={info.sa_radix_passes} "
          f"pair_passes={info.sa_pair_passes} ({info.sa_pair_passes / n:.2f} n) build={dt:.2f} s")
    assert info.sa_rounds >= 10
    assert info.sa_pair_passes < 12 * n     # round 0 sorts 8 x n pair-passes; all later rounds together stay below 4 n
    idx.close()
