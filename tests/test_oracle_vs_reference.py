"""CPU: differential test of the oracle against the compiled, unmodified reference
(oracle/_ref/libcsref.so). Skipped when the prebuilt library is absent."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/libcsref.so not built")


def _rand_text(rng, n, sigma, term):
    alpha = np.sort(rng.choice(np.arange(1, 256), sigma, replace=False)).astype(np.uint8) if sigma < 256 \
        else np.arange(256, dtype=np.uint8)
    body = alpha[rng.integers(0, sigma, n)].astype(np.uint8)
    return (np.concatenate([body, np.zeros(1, np.uint8)]) if term else body), alpha


@pytest.mark.parametrize("sigma,n", [(1, 50), (2, 999), (3, 4000), (4, 12000), (26, 6000), (256, 9000)])
def test_sa_matches_build_sa_naive(sigma, n):
    rng = np.random.default_rng(n * 7 + sigma)
    for term in (True, False):
        text, _ = _rand_text(rng, n, sigma, term)
        sa = oracle.sa_build(text)
        assert (sa == oracle.ref_sa_naive(text)).all()
        assert oracle.sa_check(text, sa) == 0
        bad = sa.copy()
        if n > 2:
            bad[[1, 2]] = bad[[2, 1]]
            assert oracle.sa_check(text, bad) != 0


def test_sa_repetitive_texts():
    for text in [b"a" * 300, b"ab" * 200, b"abc" * 100 + b"ab", b"\x00" * 64, b"\xff\x00" * 50, bytes(range(256)) * 3]:
        assert (oracle.sa_build(text) == oracle.ref_sa_naive(text)).all()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_build_products_and_queries(seed):
    rng = np.random.default_rng(seed)
    for sigma, n, stride, term in [(4, 3000, 32, True), (4, 3000, 3, False), (256, 5000, 16, True), (2, 2500, 7, True)]:
        text, alpha = _rand_text(rng, n, sigma, term)
        R = oracle.RefIndex(text.tobytes(), stride=stride)
        O = oracle.OracleIndex(text.tobytes(), stride=stride)
        assert (R.sa == O.sa).all() and (R.bwt == O.bwt).all() and (R.C == O.C).all() and (R.ssa == O.ssa).all()
        # injected-SA path equals the verbatim build (justifies csref_inject for large n)
        R2 = oracle.RefIndex(text.tobytes(), stride=stride, sa=O.sa)
        assert (R2.bwt == R.bwt).all() and (R2.ssa == R.ssa).all() and (R2.C == R.C).all()
        R3 = oracle.RefIndex(bwt=O.bwt, ssa=O.ssa, stride=stride)
        for _ in range(60):
            m = int(rng.integers(1, 12))
            if rng.random() < 0.7:
                s = int(rng.integers(0, len(text) - m))
                pat = text[s:s + m].tobytes()
            else:
                pat = alpha[rng.integers(0, sigma, m)].astype(np.uint8).tobytes()
            c = R.count(pat)
            assert O.count(pat) == c == R2.count(pat) == R3.count(pat)
            for lim in (100000, 3):
                assert O.locate(pat, lim) == R.locate(pat, lim) == R3.locate(pat, lim)
        for _ in range(200):
            c, i = int(rng.integers(0, 256)), int(rng.integers(0, len(text) + 2))
            assert O.occ(c, i) == R.occ(c, i)
        for i in rng.integers(0, len(text), 100):
            assert O.LF(int(i)) == R.LF(int(i))


def test_wavelet_directory_identical():
    rng = np.random.default_rng(5)
    seq = rng.integers(0, 256, 7001, dtype=np.uint8)
    W, V = oracle.RefWavelet(seq), oracle.OracleWavelet(seq)
    for level in range(8):
        for a, b in zip(W.level_arrays(level), V.level_arrays(level)):
            assert a.shape == b.shape and (a == b).all()
    for c in rng.integers(0, 256, 24):
        for i in list(rng.integers(0, 7003, 24)) + [0, 7001, 7002]:
            assert W.rank(int(c), int(i)) == V.rank(int(c), int(i))
    for i in rng.integers(0, 7001, 200):
        assert W.access(int(i)) == V.access(int(i)) == seq[int(i)]


def test_bitvector_every_position():
    rng = np.random.default_rng(6)
    for n in (1, 63, 64, 65, 255, 256, 257, 2047, 2048, 2049, 4500):
        for dens in (0.02, 0.5, 0.97):
            bits = (rng.random(n) < dens).astype(np.uint8)
            A, B = oracle.RefBitVector(bits), oracle.OracleBitVector(bits)
            for i in range(n + 3):
                assert A.rank1(i) == B.rank1(i)
                assert A.get(i) == B.get(i)


def test_batch_threads_equal_single():
    rng = np.random.default_rng(8)
    text, alpha = _rand_text(rng, 20000, 4, True)
    O = oracle.OracleIndex(text.tobytes(), stride=8)
    R = oracle.RefIndex(text.tobytes(), stride=8, sa=O.sa)
    pats = [text[s:s + m].tobytes() for s, m in zip(rng.integers(0, 19000, 500), rng.integers(1, 14, 500))] + [b""]
    d, o = oracle.pack_patterns(pats)
    c1, se = O.count_batch(d, o, nthreads=1)
    c4, se4 = O.count_batch(d, o, nthreads=4)
    assert (c1 == c4).all() and (se == se4).all()
    assert (R.count_batch(d, o, nthreads=4) == c1).all()
    offs, pos, st, _ = O.locate_batch(d, o, limit=50, nthreads=4)
    tot, rn, rpos = R.locate_batch(d, o, limit=50, nthreads=4, keep_positions=True)
    assert tot == offs[-1]
    for q in range(len(pats)):
        k = int(offs[q + 1] - offs[q])
        assert rn[q] == k
        assert (rpos[q * 50:q * 50 + k] == pos[int(offs[q]):int(offs[q + 1])]).all()


# ---- property test: arbitrary small texts, with and without a terminator ------------------------
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

_texts = st.one_of(
    st.binary(min_size=1, max_size=200),
    st.lists(st.sampled_from(list(b"ACGT")), min_size=1, max_size=300).map(bytes),
    st.lists(st.sampled_from([0, 1, 255]), min_size=1, max_size=120).map(bytes),
    st.builds(lambda unit, reps, tail: unit * reps + tail, st.binary(min_size=1, max_size=4), st.integers(1, 60),
              st.binary(max_size=3)),
)


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(text=_texts, stride=st.sampled_from([1, 2, 3, 4, 7, 32]), limit=st.sampled_from([1, 2, 5, 100000]),
       seed=st.integers(0, 2**31 - 1), terminate=st.booleans())
def test_oracle_equals_reference_on_arbitrary_texts(text, stride, limit, seed, terminate):
    """The restatement against the compiled reference where the reference's behaviour is odd but
    deterministic (SURVEY §8a notes 3-6): cyclic over-count without a terminator, LF walks that never
    reach a sampled row (the reference throws -> status 1), byte 0x00 as an ordinary symbol."""
    if terminate:
        text = text + b"\x00"
    R = oracle.RefIndex(text, stride=stride)
    O = oracle.OracleIndex(text, stride=stride)
    assert (R.sa == O.sa).all() and (R.bwt == O.bwt).all() and (R.C == O.C).all() and (R.ssa == O.ssa).all()
    rng = np.random.default_rng(seed)
    t = np.frombuffer(text, np.uint8)
    pats = [b"", text[-min(len(text), 30):], text[:1]]
    for _ in range(25):
        m = int(rng.integers(1, 9))
        if rng.random() < 0.7 and t.size > m:
            s = int(rng.integers(0, t.size - m + 1))
            pats.append(t[s:s + m].tobytes())
        else:
            pats.append(rng.integers(0, 256, m, dtype=np.uint8).tobytes())
    for pat in pats:
        assert O.count(pat) == R.count(pat)
        opos, ost = O.locate(pat, limit)
        rpos, rst = R.locate(pat, limit)
        assert ost == rst
        if rst == 0:
            assert opos == rpos
