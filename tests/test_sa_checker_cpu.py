"""CPU: the independent SA-binary-search checker (tests/sa_checker.py) against the oracle and the compiled
reference on small terminated texts, so that the full-size GPU tests can trust it at n = 2^28 .. 4e9."""
import numpy as np
import pytest
import torch

import oracle
from sa_checker import DeviceArrayU32, expected_locate, sa_intervals


def _text(rng, n, sigma):
    body = rng.integers(1, sigma + 1, n).astype(np.uint8)
    return np.concatenate([body, np.zeros(1, np.uint8)])  # unique smallest terminator


def _patterns(rng, text, npat, sigma, maxlen):
    pats = []
    for _ in range(npat):
        m = int(rng.integers(1, maxlen + 1))
        if rng.random() < 0.7:
            s = int(rng.integers(0, len(text) - m))
            pats.append(text[s:s + m].tobytes())
        else:
            pats.append(rng.integers(1, sigma + 1, m).astype(np.uint8).tobytes())
    pats += [b"", bytes([sigma + 1]), text[-5:-1].tobytes(), text[:7].tobytes()]
    return oracle.pack_patterns(pats)


@pytest.mark.parametrize("sigma,n,maxlen", [(2, 3000, 24), (4, 20000, 16), (200, 30000, 6), (5, 777, 40)])
def test_checker_equals_oracle(sigma, n, maxlen):
    rng = np.random.default_rng(sigma * 1000 + n)
    text = _text(rng, n, sigma)
    O = oracle.OracleIndex(text.tobytes(), stride=4)
    data, offs = _patterns(rng, text, 400, sigma, maxlen)
    t = torch.from_numpy(text)
    sa = DeviceArrayU32(torch.from_numpy(O.sa.astype(np.uint32).view(np.int32)))
    d = torch.from_numpy(data) if data.size else torch.zeros(1, dtype=torch.uint8)
    o = torch.from_numpy(offs.astype(np.int64))
    lb, ub = sa_intervals(t, sa, d, o, chunk=128)
    counts, sp_ep = O.count_batch(data, offs)
    lens = (o[1:] - o[:-1]).numpy()
    want = np.where(lens == 0, n + 1, counts.astype(np.int64))  # count("") == n (fm_index.cpp:80)
    got = (ub - lb).numpy()
    assert (got == want).all()
    hit = (counts > 0) & (lens > 0)
    assert (lb.numpy()[hit] == sp_ep[hit, 0].astype(np.int64)).all()
    assert (ub.numpy()[hit] == sp_ep[hit, 1].astype(np.int64)).all()
    for limit in (100000, 3):
        e_offs, e_pos = expected_locate(sa, lb, ub, limit, lens=o[1:] - o[:-1])
        o_offs, o_pos, o_status, _ = O.locate_batch(data, offs, limit=limit)
        assert (o_status == 0).all()
        assert (e_offs.numpy() == o_offs.astype(np.int64)).all()
        assert (e_pos.numpy() == o_pos.astype(np.int64)).all()


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/libcsref.so not built")
def test_checker_equals_compiled_reference():
    rng = np.random.default_rng(11)
    text = _text(rng, 5000, 4)
    R = oracle.RefIndex(text.tobytes(), stride=8)
    data, offs = _patterns(rng, text, 120, 4, 12)
    t = torch.from_numpy(text)
    sa = DeviceArrayU32(torch.from_numpy(R.sa.astype(np.uint32).view(np.int32)))
    lb, ub = sa_intervals(t, sa, torch.from_numpy(data), torch.from_numpy(offs.astype(np.int64)))
    for q in range(offs.size - 1):
        pat = data[int(offs[q]): int(offs[q + 1])].tobytes()
        if not pat:
            continue
        assert int(ub[q] - lb[q]) == R.count(pat)
        assert R.locate(pat, 5) == (R.sa[int(lb[q]): int(lb[q]) + min(5, int(ub[q] - lb[q]))].tolist(), 0)
