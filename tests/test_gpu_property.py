"""GPU property tests (hypothesis): arbitrary small texts over arbitrary alphabets, with and without
a terminator, arbitrary strides and limits — the engine must agree with the oracle bit for bit,
including the reference's quirks (cyclic over-count, non-terminating LF walks -> status 1)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fm():
    import csfm_b200
    csfm_b200.lib()
    return csfm_b200


texts = st.one_of(
    st.binary(min_size=1, max_size=600),
    st.lists(st.sampled_from(list(b"ACGT")), min_size=1, max_size=900).map(bytes),
    st.lists(st.sampled_from([0, 1, 255]), min_size=1, max_size=300).map(bytes),
    st.builds(lambda unit, reps, tail: unit * reps + tail, st.binary(min_size=1, max_size=5), st.integers(1, 150),
              st.binary(max_size=3)),
)


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(text=texts, stride=st.sampled_from([1, 2, 3, 4, 7, 32, 1000]), limit=st.sampled_from([1, 2, 5, 100000]),
       flags=st.sampled_from([0, 1, 4, 5, 8, 32, 33, 40, 64, 65, 128, 136]), seed=st.integers(0, 2**31 - 1), terminate=st.booleans())
def test_engine_equals_oracle(fm, text, stride, limit, flags, seed, terminate):
    if terminate:
        text = text + b"\x00"
    t = np.frombuffer(text, np.uint8)
    idx = fm.FMIndex.build_from_text(t, fm.BuildParams(ssa_stride=stride), flags=flags | fm.BUILD_KEEP_SA)
    orc = oracle.OracleIndex(t, stride=stride)
    assert (idx.sa() == orc.sa).all()
    assert (idx.bwt() == orc.bwt).all() and (idx.ssa() == orc.ssa).all() and (idx.C_array() == orc.C).all()
    rng = np.random.default_rng(seed)
    pats = [b""]
    for _ in range(40):
        m = int(rng.integers(1, 9))
        if rng.random() < 0.7 and t.size > m:
            s = int(rng.integers(0, t.size - m + 1))
            pats.append(t[s:s + m].tobytes())
        else:
            pats.append(rng.integers(0, 256, m, dtype=np.uint8).tobytes())
    pats.append(text[-min(len(text), 40):])
    d, o = fm.pack_patterns(pats)
    counts, sp_ep = idx.count_batch(d, o, want_intervals=True)
    oc, ose = orc.count_batch(d, o)
    assert (counts == oc).all() and (sp_ep == ose).all()
    assert (idx.count_batch(d, o) == oc).all()  # without intervals: may take the text-verification shortcut
    offs, pos, status = idx.locate_batch(d, o, limit=limit)
    ooffs, opos, ostatus, _ = orc.locate_batch(d, o, limit=limit)
    assert (offs == ooffs).all() and (status == ostatus).all()
    ok = np.repeat(ostatus == 0, np.diff(ooffs).astype(np.int64))
    assert (pos[ok] == opos[ok]).all()
    # the one-pattern calls of the reference's API (single-launch path on layout 2, batch path elsewhere)
    for q in (1, len(pats) // 2, len(pats) - 1):
        assert idx.count(pats[q]) == int(oc[q])
        if ostatus[q] == 0:
            assert idx.locate(pats[q], limit) == opos[int(ooffs[q]): int(ooffs[q + 1])].tolist()
    # the packed wire form of the streaming API (patterns of present symbols only)
    codes, bits = idx.pattern_codes()
    keep = [p for p in pats if all(codes[b] != 255 for b in p)]
    dk, ok_ = fm.pack_patterns(keep)
    want = orc.count_batch(dk, ok_)[0] if keep else np.zeros(0, np.uint64)
    packed = idx.pack_codes(dk) if dk.size else np.zeros(1, np.uint8)
    lens = np.diff(ok_).astype(np.uint8)
    out = np.zeros(max(1, len(keep)), np.uint32)
    idx.count_batch_wait(idx.count_batch_submit_packed(packed.ctypes.data, int(dk.size), lens.ctypes.data, len(keep), out.ctypes.data))
    assert (out[: len(keep)].astype(np.uint64) == want).all()
    # extract from the index itself (no host text behind a handle made from the blob): texts with a unique smallest terminator
    if terminate and (t[:-1] != 0).all() and not (flags & 4):
        lean = fm.FMIndex.from_host_blob(idx.blob_to_host())
        assert lean.extract(0, t.size + 3) == text and lean.extract(t.size // 2, 5) == text[t.size // 2: t.size // 2 + 5]
