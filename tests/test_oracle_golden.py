"""CPU: the plain-C oracle (oracle/fm_oracle.c) against the committed golden vectors.

The vectors were produced by the UNMODIFIED reference (tests/golden/make_golden.py), so this
pins the oracle on machines where /root/reference does not exist (the GPU box).
"""
import json
import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FM = json.load(open(os.path.join(GOLDEN, "golden_fm.json")))
WV = json.load(open(os.path.join(GOLDEN, "golden_wavelet.json")))


@pytest.mark.parametrize("case", FM["cases"], ids=[c["name"] for c in FM["cases"]])
def test_fm_case(case):
    text = bytes.fromhex(case["text_hex"])
    O = oracle.OracleIndex(text, stride=case["stride"])
    assert O.n == case["n"]
    assert O.bwt.tobytes() == bytes.fromhex(case["bwt_hex"])
    assert O.C.tolist() == case["C"]
    assert O.ssa.tolist() == case["ssa"]
    if "sa" in case:
        assert O.sa.tolist() == case["sa"]
    assert oracle.sa_check(text, O.sa) == 0
    for q in case["queries"]:
        pat = bytes.fromhex(q["pat_hex"])
        assert O.count(pat) == q["count"], (case["name"], pat)
        pos, st = O.locate(pat, q["limit"])
        assert st == q["status"], (case["name"], pat)
        assert pos == q["locate"], (case["name"], pat, q["limit"])


@pytest.mark.parametrize("case", WV["wavelet"], ids=[c["name"] for c in WV["wavelet"]])
def test_wavelet_case(case):
    seq = bytes.fromhex(case["seq_hex"])
    W = oracle.OracleWavelet(seq)
    for c, i, r in case["ranks"]:
        assert W.rank(c, i) == r, (case["name"], c, i)
    for i, a in enumerate(case["access_prefix"]):
        assert W.access(i) == a


@pytest.mark.parametrize("case", WV["bitvector"], ids=[c["name"] for c in WV["bitvector"]])
def test_bitvector_case(case):
    if "word" in case:
        B = oracle.OracleBitVector(words=np.full(case["nwords"], case["word"], dtype=np.uint64), nbits=case["nbits"])
    else:
        packed = np.frombuffer(bytes.fromhex(case["bits_packed_hex"]), dtype=np.uint8)
        bits = np.unpackbits(packed, bitorder="little")[: case["nbits"]]
        B = oracle.OracleBitVector(bits)
    for i, r in zip(case["positions"], case["rank1"]):
        assert B.rank1(i) == r, (case["name"], i)


def test_c1_benchmark_workload():
    """BASELINE.json configs[0]: the reference's own benchmark inputs and checksums."""
    z = np.load(os.path.join(GOLDEN, "c1_workload.npz"))
    text = z["text"]
    O = oracle.OracleIndex(text, stride=32)
    assert (O.sa == z["sa"]).all()
    assert (O.ssa == z["ssa"]).all()
    assert (O.C == z["C"]).all()
    pats = [text[p:p + 5].tobytes() for p in z["rand_pos"]]
    d, o = oracle.pack_patterns(pats)
    counts, _ = O.count_batch(d, o)
    assert (counts == z["rand_count"]).all()
    assert int(counts.sum()) == 9907582  # total_matches of the unmodified benchmark
    freq = [bytes(p)[:l] for p, l in zip(z["freq_patterns"], z["freq_len"])]
    d, o = oracle.pack_patterns(freq)
    fcounts, _ = O.count_batch(d, o)
    assert (fcounts == z["freq_count"]).all()
    assert int(fcounts.sum()) * 1000 == 16309000
    offs, pos, status, _ = O.locate_batch(d, o, limit=100000)
    assert (status == 0).all()
    for q in range(len(freq)):
        a = pos[int(offs[q]):int(offs[q + 1])]
        assert a.size == z["loc_n"][q]
        assert int(a.sum()) == int(z["loc_sum"][q])
        assert int(np.bitwise_xor.reduce(a)) == int(z["loc_xor"][q])
        assert a[:8].tolist() == z["loc_first"][q][: a.size].tolist()
    assert int(offs[-1]) * 10 == 163090
