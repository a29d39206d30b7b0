"""GPU parity: the CUDA engine (through the C ABI, include/csfm.h) against
 (a) the committed golden vectors produced by the unmodified reference, and
 (b) the CPU oracle on seeded inputs, bit-exact: counts, [sp,ep), positions in SA-row order,
     per-query status, and the build products (SA, BWT, C, SSA)."""
import json
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FM = json.load(open(os.path.join(GOLDEN, "golden_fm.json")))


@pytest.fixture(scope="module")
def fm():
    import csfm_b200
    csfm_b200.lib()  # raises if libcsfm.so is missing: there is no fallback to test instead
    return csfm_b200


def _expected_lookups(info, text, d, o):
    """Queries that start from the k-mer table: length >= k; on layout 3 the keys are made of the two-bit codes,
    so a pattern whose last k bytes contain the symbol that occurs once starts from the C array instead."""
    kk = int(info.kmer_k)
    lens = np.diff(o).astype(np.int64)
    use = lens >= kk
    hist = np.bincount(np.asarray(text, dtype=np.uint8), minlength=256)
    if info.layout == 3 and (hist > 0).sum() == 5:
        single = int(np.flatnonzero(hist == 1)[0])
        ends = o[1:].astype(np.int64)
        for q in np.flatnonzero(use):
            if (d[ends[q] - kk: ends[q]] == single).any():
                use[q] = False
    return int(use.sum())


def _check_queries(fm, idx, case):
    pats = [bytes.fromhex(q["pat_hex"]) for q in case["queries"]]
    d, o = fm.pack_patterns(pats)
    counts = idx.count_batch(d, o)
    for q, c in zip(case["queries"], counts):
        assert int(c) == q["count"], (case["name"], q["pat_hex"])
    for lim in sorted({q["limit"] for q in case["queries"]}):
        sel = [q for q in case["queries"] if q["limit"] == lim]
        d, o = fm.pack_patterns([bytes.fromhex(q["pat_hex"]) for q in sel])
        offs, pos, status = idx.locate_batch(d, o, limit=lim)
        for k, q in enumerate(sel):
            assert int(status[k]) == q["status"], (case["name"], q["pat_hex"], lim)
            if q["status"] == 0:
                got = pos[int(offs[k]):int(offs[k + 1])].tolist()
                assert got == q["locate"], (case["name"], q["pat_hex"], lim)
    # single-query API mirrors the reference's exceptions
    for q in case["queries"][:6]:
        pat = bytes.fromhex(q["pat_hex"])
        assert idx.count(pat) == q["count"]
        if q["status"] == 0:
            assert idx.locate(pat, q["limit"]) == q["locate"]
        else:
            with pytest.raises(RuntimeError, match="LF walk exceeded text length"):
                idx.locate(pat, q["limit"])


@pytest.mark.parametrize("case", FM["cases"], ids=[c["name"] for c in FM["cases"]])
@pytest.mark.parametrize("flags", [0, 128, 1, 4, 5, 32], ids=["default", "nib128", "nib128-rawbytes", "bin64", "bin64-8levels", "nib128-textcheck"])
def test_golden_from_text(fm, case, flags):
    """build_from_text on the GPU (SA -> BWT -> C -> wavelet -> SSA) + queries vs the reference."""
    text = bytes.fromhex(case["text_hex"])
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=case["stride"]), flags=flags | fm.BUILD_KEEP_SA)
    info = idx.info()
    assert info.n == case["n"]
    hist = np.bincount(np.frombuffer(text, np.uint8), minlength=256)
    dna_ok = len(text) > 0 and ((hist > 0).sum() <= 4 or ((hist > 0).sum() == 5 and (hist == 1).any()))
    want_layout = 1 if flags & 4 else (3 if flags == 0 and dna_ok else 2)   # layout 3 is the default where the text allows it
    assert info.layout == want_layout and info.line_bytes == (128 if want_layout == 2 else 64)
    assert info.levels == {0: info.levels, 128: info.levels, 1: 2, 4: info.levels, 5: 8, 32: info.levels}[flags] and 1 <= info.levels <= 8
    assert idx.C_array().tolist() == case["C"]
    assert idx.ssa().tolist() == case["ssa"]
    assert idx.bwt().tobytes() == bytes.fromhex(case["bwt_hex"])
    if "sa" in case:
        assert idx.sa().tolist() == case["sa"]
    _check_queries(fm, idx, case)
    assert idx.extract(1, 3) == text[1:4]


@pytest.mark.parametrize("case", FM["cases"][:12], ids=[c["name"] for c in FM["cases"][:12]])
def test_golden_from_parts(fm, case):
    """csfm_build_from_parts: reference BWT + SSA in, same answers out."""
    if case["n"] == 0:
        pytest.skip("nothing to inject")
    idx = fm.FMIndex.from_parts(bytes.fromhex(case["bwt_hex"]), np.array(case["ssa"], np.uint32), case["stride"])
    _check_queries(fm, idx, case)


def test_c1_benchmark_workload(fm):
    """BASELINE.json configs[0] (tools/benchmark.cpp) with the reference's total_matches checksums."""
    z = np.load(os.path.join(GOLDEN, "c1_workload.npz"))
    text = z["text"]
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(), flags=fm.BUILD_KEEP_SA)
    assert (idx.sa() == z["sa"]).all()
    assert (idx.ssa() == z["ssa"]).all()
    assert (idx.C_array() == z["C"]).all()
    d, o = fm.pack_patterns([text[p:p + 5].tobytes() for p in z["rand_pos"]])
    counts = idx.count_batch(d, o)
    assert (counts == z["rand_count"]).all()
    assert int(counts.sum()) == 9907582
    freq = [bytes(p)[:l] for p, l in zip(z["freq_patterns"], z["freq_len"])]
    d, o = fm.pack_patterns([freq[i % 10] for i in range(10000)])
    fcounts = idx.count_batch(d, o)
    assert int(fcounts.sum()) == 16309000
    d, o = fm.pack_patterns([freq[i % 10] for i in range(100)])
    offs, pos, status = idx.locate_batch(d, o, limit=100000)
    assert (status == 0).all() and int(offs[-1]) == 163090
    for q in range(100):
        a = pos[int(offs[q]):int(offs[q + 1])]
        assert a.size == z["loc_n"][q % 10]
        assert int(a.sum()) == int(z["loc_sum"][q % 10])
        assert int(np.bitwise_xor.reduce(a)) == int(z["loc_xor"][q % 10])
        assert a[:8].tolist() == z["loc_first"][q % 10][: a.size].tolist()


def _rand_text(rng, n, sigma, term):
    alpha = np.sort(rng.choice(np.arange(1, 256), sigma, replace=False)).astype(np.uint8) if sigma < 256 \
        else np.arange(256, dtype=np.uint8)
    body = alpha[rng.integers(0, sigma, n)].astype(np.uint8)
    return (np.concatenate([body, np.zeros(1, np.uint8)]) if term else body), alpha


def _mixed_patterns(rng, text, alpha, k, maxlen):
    pats = []
    for _ in range(k):
        m = int(rng.integers(1, maxlen + 1))
        r = rng.random()
        if r < 0.6 and text.size > m:
            s = int(rng.integers(0, text.size - m))
            pats.append(text[s:s + m].tobytes())
        elif r < 0.9:
            pats.append(alpha[rng.integers(0, alpha.size, m)].astype(np.uint8).tobytes())
        else:
            pats.append(rng.integers(0, 256, m, dtype=np.uint8).tobytes())  # bytes outside the alphabet
    pats += [b"", text[-3:].tobytes(), text[:1].tobytes()]
    return pats


@pytest.mark.parametrize("layout", [0, 4, 128], ids=["default", "bin64", "nib128"])
@pytest.mark.parametrize("sigma,n,stride,term,flags", [
    (2, 50_000, 4, True, 0), (4, 200_000, 32, True, 0), (4, 200_000, 32, True, 1), (5, 100_000, 7, False, 0),
    (15, 90_000, 5, True, 0), (16, 90_000, 5, False, 0), (17, 90_000, 5, True, 0),
    (21, 150_000, 16, True, 0), (97, 120_000, 32, True, 0), (256, 300_000, 32, True, 0), (256, 65_537, 1, False, 0),
    (1, 5_000, 3, True, 0), (3, 479, 2, True, 0), (3, 480, 2, True, 0), (3, 481, 2, False, 0), (3, 961, 5, True, 0),
    (3, 127, 2, True, 0), (3, 128, 2, False, 0), (3, 129, 2, True, 0), (40, 255, 3, True, 0), (40, 256, 3, False, 0),
])
def test_random_vs_oracle(fm, sigma, n, stride, term, flags, layout):
    flags |= layout
    rng = np.random.default_rng(1000 * sigma + n + stride)
    text, alpha = _rand_text(rng, n, sigma, term)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=stride), flags=flags | fm.BUILD_KEEP_SA)
    orc = oracle.OracleIndex(text, stride=stride)
    assert (idx.sa() == orc.sa).all()
    assert (idx.bwt() == orc.bwt).all()
    assert (idx.C_array() == orc.C).all()
    assert (idx.ssa() == orc.ssa).all()
    pats = _mixed_patterns(rng, text, alpha, 3000, 24)
    d, o = fm.pack_patterns(pats)
    oc, ose, osteps = orc.count_batch(d, o, want_steps=True)
    idx.set_instrumentation(1)
    counts, sp_ep = idx.count_batch(d, o, want_intervals=True)
    st = idx.last_call_stats()
    assert (counts == oc).all()
    assert (sp_ep == ose).all()
    kk = idx.info().kmer_k
    if kk == 0:
        assert st.search_steps == int(osteps.sum())  # the S of the roofline model is counted exactly
    else:
        # queries of length >= k start from the k-mer jump table: same answers, fewer rank steps
        assert st.table_lookups == _expected_lookups(idx.info(), text, d, o)
        assert st.search_steps <= int(osteps.sum())
        plain = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=stride), flags=flags | fm.BUILD_NO_KMER_TABLE)
        assert plain.info().kmer_k == 0
        plain.set_instrumentation(1)
        c2, se2 = plain.count_batch(d, o, want_intervals=True)
        assert (c2 == oc).all() and (se2 == ose).all()
        assert plain.last_call_stats().search_steps == int(osteps.sum())
    for limit in (100000, 7):
        offs, pos, status = idx.locate_batch(d, o, limit=limit)
        lf_gpu = idx.last_call_stats().lf_steps
        ooffs, opos, ostatus, olf = orc.locate_batch(d, o, limit=limit)
        assert (offs == ooffs).all()
        assert (status == ostatus).all()
        ok = np.repeat(ostatus == 0, np.diff(ooffs).astype(np.int64))
        assert (pos[ok] == opos[ok]).all()
        if idx.info().position_samples:  # layout 3, marked form: a walk takes SA[row] mod stride steps
            assert (ostatus == 0).all() and lf_gpu == int((opos % stride).sum())
        elif (ostatus == 0).all():
            assert lf_gpu == olf


@pytest.mark.parametrize("letters,n,stride,where,single", [
    (4, 150_000, 32, "end", 0x24), (4, 150_000, 8, "middle", 0x24), (4, 40_000, 5, "start", 0xFF), (4, 40_000, 3, "middle", 0x42),
    (3, 1_000, 4, "end", 0x00), (4, 191, 2, "end", 0x01), (4, 192, 2, "middle", 0x01), (4, 193, 2, "start", 0x01),
    (4, 383, 7, "middle", 0x7F), (4, 384, 7, "end", 0x7F), (4, 385, 1, "start", 0x7F), (2, 5_000, 16, None, 0), (4, 97, 32, None, 0),
    (1, 700, 9, "end", 0x00),
])
@pytest.mark.parametrize("lanes", ["1", "2"])
def test_dna_layout_vs_oracle(fm, monkeypatch, lanes, letters, n, stride, where, single):
    """Layout 3 (two-bit symbols, 64-byte lines): <= 4 frequent symbols plus one symbol that
    occurs once ANYWHERE in the text and anywhere in byte order (not only a smallest terminator), line-boundary
    sizes (192 symbols per line), failing LF walks, every build product, counts, intervals, positions. Both forms of
    the count and walk kernels: one lane per query / row (the default while the index fits the L2) and a two-lane
    sub-warp (the default beyond); the knobs are read when a handle is created."""
    monkeypatch.setenv("CSFM_COUNT3_LANES", lanes)
    monkeypatch.setenv("CSFM_WALK3_LANES", lanes)
    rng = np.random.default_rng(letters * 7919 + n + stride)
    alpha = np.array([0x41, 0x43, 0x47, 0x54][:letters], dtype=np.uint8)
    body = alpha[rng.integers(0, letters, n)].astype(np.uint8)
    if where is not None:
        at = {"end": n - 1, "middle": n // 2, "start": 0}[where]
        body[at] = single
    text = body
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=stride), flags=fm.BUILD_KEEP_SA)
    info = idx.info()
    assert info.layout == 3 and info.line_bytes == 64 and info.levels == 1
    # the marked line form (suffix array sampled by text position) exactly when the last byte is the single symbol
    assert info.position_samples == (1 if where == "end" else 0)
    assert info.blocks_per_level == n // (128 if info.position_samples else 192) + 1
    orc = oracle.OracleIndex(text, stride=stride)
    assert (idx.sa() == orc.sa).all()
    assert (idx.bwt() == orc.bwt).all()          # access through the two-bit lines, incl. the row of the single symbol
    assert (idx.C_array() == orc.C).all()
    assert (idx.ssa() == orc.ssa).all()
    both = np.concatenate([alpha, np.array([single], np.uint8)]) if where is not None else alpha
    pats = _mixed_patterns(rng, text, both, 4000, 30)
    if where is not None:  # patterns through, ending at and starting at the single symbol
        at = int(np.flatnonzero(text == single)[0])
        for a, b in ((at - 3, at + 4), (at - 5, at + 1), (at, at + 6), (at, at + 1), (at - 1, at + 1)):
            a, b = max(a, 0), min(b, n)
            pats.append(text[a:b].tobytes())
        pats.append(bytes([single, single]))
    d, o = fm.pack_patterns(pats)
    oc, ose, osteps = orc.count_batch(d, o, want_steps=True)
    rows_only = None
    for build_flags in (0, fm.BUILD_NO_KMER_TABLE, fm.BUILD_ROW_SAMPLES):
        ix = idx if build_flags == 0 else fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=stride), flags=build_flags)
        assert ix.info().layout == 3
        if build_flags == fm.BUILD_ROW_SAMPLES:
            rows_only = ix
            assert ix.info().position_samples == 0 and (ix.ssa() == orc.ssa).all() and (ix.bwt() == orc.bwt).all()
        ix.set_instrumentation(1)
        counts, sp_ep = ix.count_batch(d, o, want_intervals=True)
        st = ix.last_call_stats()
        assert (counts == oc).all()
        assert (sp_ep == ose).all()
        if ix.info().kmer_k == 0:
            assert st.search_steps == int(osteps.sum())
        else:
            assert st.table_lookups == _expected_lookups(ix.info(), text, d, o)
        ix.set_instrumentation(0)
        assert (ix.count_batch(d, o) == oc).all()   # the uninstrumented kernel
    for limit in (100000, 5):
        ooffs, opos, ostatus, olf = orc.locate_batch(d, o, limit=limit)
        for ix in (idx, rows_only):
            ix.set_instrumentation(1)
            offs, pos, status = ix.locate_batch(d, o, limit=limit)
            lf_gpu = ix.last_call_stats().lf_steps
            assert (offs == ooffs).all()
            assert (status == ostatus).all()
            ok = np.repeat(ostatus == 0, np.diff(ooffs).astype(np.int64))
            assert (pos[ok] == opos[ok]).all()
            if ix.info().position_samples:
                # a walk ends at the first row whose suffix starts at a multiple of the stride: SA[row] mod stride steps
                assert (ostatus == 0).all() and lf_gpu == int((opos % stride).sum())
            elif (ostatus == 0).all():
                assert lf_gpu == olf
    # the same answers from layout 2 on the same text, from a blob round trip, and one query at a time
    idx2 = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=stride), flags=fm.BUILD_LAYOUT_NIBBLE128)
    assert idx2.info().layout == 2
    assert (idx2.count_batch(d, o) == oc).all()
    back = fm.FMIndex.from_host_blob(idx.blob_to_host())
    assert back.info().layout == 3
    c3, se3 = back.count_batch(d, o, want_intervals=True)
    assert (c3 == oc).all() and (se3 == ose).all()
    assert (back.bwt() == orc.bwt).all()
    assert back.info().position_samples == info.position_samples      # the attached blob walks like the one it was copied from
    ooffs, opos, ostatus, _ = orc.locate_batch(d, o, limit=9)
    offs, pos, status = back.locate_batch(d, o, limit=9)
    ok = np.repeat(ostatus == 0, np.diff(ooffs).astype(np.int64))
    assert (offs == ooffs).all() and (status == ostatus).all() and (pos[ok] == opos[ok]).all()
    for q in range(0, len(pats), max(1, len(pats) // 40)):
        assert idx.count(pats[q]) == int(oc[q])


@pytest.mark.parametrize("letters,n,stride", [(4, 100_000, 32), (4, 128 * 37, 8), (4, 128 * 37 - 1, 3), (4, 129, 1), (3, 5_000, 7), (1, 300, 4)])
def test_marked_lines_build_products(fm, letters, n, stride):
    """The marked line form of layout 3 read back word by word (csrc/csfm_dna.cuh): 128 rows per 64-byte line, a mark bit
    exactly at the rows whose suffix starts at a multiple of the stride, three counters + the mark counter before the
    line, and the position samples = SA values of the marked rows in row order — against the oracle's SA and BWT."""
    rng = np.random.default_rng(n + stride)
    alpha = np.array([0x41, 0x43, 0x47, 0x54][:letters], dtype=np.uint8)
    text = np.concatenate([alpha[rng.integers(0, letters, n - 1)], np.array([0x24], np.uint8)]).astype(np.uint8)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=stride), flags=fm.BUILD_KEEP_SA)
    info = idx.info()
    assert info.layout == 3 and info.position_samples == 1
    orc = oracle.OracleIndex(text, stride=stride)
    sa = orc.sa.astype(np.int64)
    nblk, nsamp = n // 128 + 1, (n + stride - 1) // stride
    assert info.blocks_per_level == nblk and info.nsamp == nsamp
    blob = idx.blob_to_host()
    al = lambda x: (x + 255) // 256 * 256
    off_levels = 4096
    off_ssa = off_levels + al(nblk * 64)
    off_psamp = al(off_ssa + nsamp * 4)
    lines = blob[off_levels: off_levels + nblk * 64].view(np.uint32).reshape(nblk, 16)
    # expected symbols: compact two-bit codes in byte order; with five symbols the terminator has none and is stored as 0
    present = np.unique(text)
    coded = present if present.size <= 4 else present[present != 0x24]
    code = np.zeros(256, np.int64)
    code[coded] = np.arange(coded.size)
    from test_marked_model_cpu import marked_lines      # the numpy statement of the format (its walk is checked on the CPU)
    want_lines, want_psamp = marked_lines(code[orc.bwt], sa, n, stride)
    for w in range(16):
        assert (lines[:, w] == want_lines[:, w]).all(), w
    psamp = blob[off_psamp: off_psamp + nsamp * 4].view(np.uint32)
    assert (psamp == want_psamp).all() and (psamp == sa[sa % stride == 0]).all()          # boolean indexing keeps row order
    assert (blob[off_ssa: off_ssa + nsamp * 4].view(np.uint32) == orc.ssa).all()   # the reference's row samples stay for export
    # and every row walks to its own suffix: locate of all one-character patterns returns SA in row order
    d, o = fm.pack_patterns([bytes([b]) for b in present])
    offs, pos, status = idx.locate_batch(d, o, limit=n)
    assert (status == 0).all() and int(offs[-1]) == n
    C = orc.C.astype(np.int64)
    for q, b in enumerate(present):
        assert (pos[int(offs[q]): int(offs[q + 1])] == sa[C[b]: C[b + 1]]).all()


@pytest.mark.parametrize("sigma,n,stride,flags", [(4, 70_000, 32, 0), (4, 70_000, 5, 128), (60, 90_000, 16, 0), (200, 50_000, 7, 32),
                                                  (3, 1_000, 1, 0), (255, 300, 4, 0)])
def test_extract_from_the_index(fm, sigma, n, stride, flags):
    """cs::FMIndex::extract (fm_index.cpp:163-167) on a handle that has no host copy of the text: csfm_extract serves
    it from the blob's text section, or rebuilds the text out of the index (every sampled row walks LF to the next)."""
    rng = np.random.default_rng(sigma + n)
    text, _ = _rand_text(rng, n, sigma, True)
    built = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=stride), flags=flags)
    idx = fm.FMIndex.from_host_blob(built.blob_to_host())    # no host text behind this handle
    raw = text.tobytes()
    assert idx.extract(0, len(raw)) == raw
    for p, l in [(0, 1), (5, 17), (len(raw) - 4, 100), (len(raw) - 1, 1), (len(raw), 5), (len(raw) + 7, 1), (123, 0)]:
        assert idx.extract(p, l) == raw[p:p + l] if p < len(raw) else idx.extract(p, l) == b""
    assert built.extract(3, 9) == raw[3:12]                  # the host-text path of the reference


def test_extract_needs_a_terminated_text(fm):
    idx = fm.FMIndex.from_host_blob(fm.FMIndex.build_from_text(b"abab" * 50, fm.BuildParams(ssa_stride=7)).blob_to_host())
    with pytest.raises(fm.CsfmError, match="cannot be rebuilt"):
        idx.extract(0, 10)
    lay1 = fm.FMIndex.from_host_blob(fm.FMIndex.build_from_text(b"banana$", fm.BuildParams(ssa_stride=2),
                                                                flags=fm.BUILD_LAYOUT_BINARY64).blob_to_host())
    with pytest.raises(fm.CsfmError, match="layout 2 or 3"):
        lay1.extract(0, 3)


@pytest.mark.parametrize("sigma,n", [(4, 120_000), (2, 9_000), (1, 500), (20, 60_000), (255, 80_000)])
def test_async_submit_packed_codes(fm, sigma, n):
    """csfm_count_batch_submit_packed: patterns as packed wire codes (3 bits per symbol for DNA + terminator)."""
    rng = np.random.default_rng(sigma * 31 + n)
    text, alpha = _rand_text(rng, n, sigma, True)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=16))
    orc = oracle.OracleIndex(text, stride=16)
    codes, bits = idx.pattern_codes()
    present = np.unique(text)
    assert bits == max(1, int(np.ceil(np.log2(present.size)))) and (codes[present] == np.arange(present.size)).all()
    pats = [p for p in _mixed_patterns(rng, text, alpha, 5000, 40) if len(p) <= 255 and all(codes[b] != 255 for b in p)]
    pats += [b"", text[-3:].tobytes()]
    d, o = fm.pack_patterns(pats)
    oc, _ = orc.count_batch(d, o)
    packed = idx.pack_codes(d)
    assert packed.size == (d.size * bits + 7) // 8
    lens = np.diff(o).astype(np.uint8)
    out = np.zeros(len(pats), np.uint32)
    t = idx.count_batch_submit_packed(packed.ctypes.data, d.size, lens.ctypes.data, len(pats), out.ctypes.data)
    idx.count_batch_wait(t)
    assert (out.astype(np.uint64) == oc).all()
    # a code beyond the alphabet stands for a symbol that does not occur: count 0
    if present.size < (1 << bits):
        bad = np.packbits(((np.full(4, (1 << bits) - 1)[:, None] >> np.arange(bits)[None, :]) & 1).astype(np.uint8).reshape(-1), bitorder="little")
        out1 = np.full(1, 7, np.uint32)
        l1 = np.array([4], np.uint8)
        idx.count_batch_wait(idx.count_batch_submit_packed(bad.ctypes.data, 4, l1.ctypes.data, 1, out1.ctypes.data))
        assert out1[0] == 0
    with pytest.raises(fm.CsfmError, match="sum of the pattern lengths"):
        idx.count_batch_submit_packed(packed.ctypes.data, d.size + 1, lens.ctypes.data, len(pats), out.ctypes.data)


def test_alias_handles_from_threads(fm):
    """csfm_alias: a second handle over the same blob (no copy) per host thread; every thread gets the oracle's answers
    while the others are running, and the aliases see the same index as the handle that owns the blob."""
    import threading
    rng = np.random.default_rng(3)
    text, alpha = _rand_text(rng, 200_000, 4, True)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=16))
    orc = oracle.OracleIndex(text, stride=16)
    pats = _mixed_patterns(rng, text, alpha, 20_000, 24)
    d, o = fm.pack_patterns(pats)
    oc, ose = orc.count_batch(d, o)
    ooffs, opos, ostatus, _ = orc.locate_batch(d[: int(o[2000])], o[:2001], limit=9)
    blob_ptr, blob_bytes = idx.blob()
    errors = []

    def worker(k):
        try:
            a = idx.alias()
            assert a.blob() == (blob_ptr, blob_bytes) and a.info().layout == idx.info().layout
            for _ in range(5):
                c, se = a.count_batch(d, o, want_intervals=True)
                assert (c == oc).all() and (se == ose).all()
                offs, pos, status = a.locate_batch(d[: int(o[2000])], o[:2001], limit=9)
                assert (offs == ooffs).all() and (pos == opos).all() and (status == ostatus).all()
                assert a.count(pats[k]) == int(oc[k])
            a.close()
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in threads:
        t.start()
    for _ in range(5):
        assert (idx.count_batch(d, o) == oc).all()     # the owning handle keeps working meanwhile
    for t in threads:
        t.join()
    assert not errors, errors


def test_new_entry_points_edge_cases(fm):
    """Argument edges of the round-2 entry points: extract at and beyond the end, empty packed batches, aliases that
    outlive nothing, single-pattern calls at the 64-byte boundary of the one-launch path."""
    text = np.concatenate([np.frombuffer(b"abracadabra" * 40, np.uint8), np.zeros(1, np.uint8)])
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=4))
    orc = oracle.OracleIndex(text, stride=4)
    lean = fm.FMIndex.from_host_blob(idx.blob_to_host())
    raw = text.tobytes()
    assert lean.extract(len(raw) - 1, 1) == raw[-1:] and lean.extract(len(raw), 1) == b"" and lean.extract(0, 0) == b""
    assert lean.extract(3, 10**12) == raw[3:]
    # empty packed batch, batch of empty patterns
    out = np.zeros(4, np.uint32)
    lens0 = np.zeros(3, np.uint8)
    one = np.zeros(1, np.uint8)
    idx.count_batch_wait(idx.count_batch_submit_packed(one.ctypes.data, 0, lens0.ctypes.data, 0, out.ctypes.data))
    idx.count_batch_wait(idx.count_batch_submit_packed(one.ctypes.data, 0, lens0.ctypes.data, 3, out.ctypes.data))
    assert out[:3].tolist() == [len(raw)] * 3                       # count("") == n (fm_index.cpp:80)
    # single-pattern calls around the parameter-space limit (64 bytes) agree with the batch kernels and the oracle
    for m in (1, 2, 63, 64, 65, 200):
        pat = raw[5:5 + m]
        want = orc.count(pat)
        assert idx.count(pat) == want
        d, o = fm.pack_patterns([pat, pat])
        assert idx.count_batch(d, o).tolist() == [want, want]
        assert idx.locate(pat, 7) == orc.locate(pat, 7)[0]
    # an alias closed before its owner, and a second one taken afterwards
    a = idx.alias()
    assert a.count(b"abra") == orc.count(b"abra")
    a.close()
    b = idx.alias()
    assert b.locate(b"cad", 3) == orc.locate(b"cad", 3)[0]
    b.close()
    assert idx.count(b"abra") == orc.count(b"abra")


def test_host_offsets_are_checked(fm):
    idx = fm.FMIndex.build_from_text(b"mississippi$", fm.BuildParams(ssa_stride=4))
    d = np.frombuffer(b"ssiissi", np.uint8)
    bad = np.array([0, 5, 3, 7], np.uint64)
    with pytest.raises(fm.CsfmError, match="must not decrease"):
        idx.count_batch(d, bad)
    with pytest.raises(fm.CsfmError, match="must not decrease"):
        idx.locate_batch(d, bad)


def test_long_and_many_patterns(fm):
    rng = np.random.default_rng(99)
    text, alpha = _rand_text(rng, 400_000, 4, True)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=32))
    orc = oracle.OracleIndex(text, stride=32, sa=None)
    # very long patterns (beyond any staging buffer) and a batch much larger than the grid
    pats = [text[s:s + m].tobytes() for s, m in zip(rng.integers(0, 300_000, 64), rng.integers(500, 5000, 64))]
    pats += [text[s:s + 12].tobytes() for s in rng.integers(0, 399_000, 200_000)]
    d, o = fm.pack_patterns(pats)
    counts, sp_ep = idx.count_batch(d, o, want_intervals=True)
    oc, ose = orc.count_batch(d, o)
    assert (counts == oc).all() and (sp_ep == ose).all()
    assert (counts >= 1).all()


def test_empty_batch_and_empty_patterns(fm):
    idx = fm.FMIndex.build_from_text(b"banana$", fm.BuildParams(ssa_stride=2))
    d, o = fm.pack_patterns([])
    assert idx.count_batch(d, o).size == 0
    offs, pos, status = idx.locate_batch(d, o)
    assert offs.tolist() == [0] and pos.size == 0
    d, o = fm.pack_patterns([b"", b"", b"a", b""])
    assert idx.count_batch(d, o).tolist() == [7, 7, 3, 7]  # count("") == n (fm_index.cpp:80)
    offs, pos, status = idx.locate_batch(d, o)
    assert offs.tolist() == [0, 0, 0, 3, 3]  # locate("") is empty (fm_index.cpp:109)
    assert pos.tolist() == [5, 3, 1]


def test_locate_capacity_contract(fm):
    import ctypes as C
    idx = fm.FMIndex.build_from_text(b"abababab$", fm.BuildParams(ssa_stride=4))
    d, o = fm.pack_patterns([b"ab", b"b"])
    out_offs = np.zeros(3, np.uint64)
    pos = np.zeros(2, np.uint64)
    status = np.zeros(2, np.int32)
    total = C.c_uint64()
    rc = fm.lib().csfm_locate_batch(idx._h, d.ctypes.data, o.ctypes.data, 2, 100, out_offs.ctypes.data, pos.ctypes.data, 2,
                                    status.ctypes.data, C.byref(total))
    assert rc == 5 and total.value == 8 and out_offs.tolist() == [0, 4, 8]  # CSFM_ERR_CAPACITY


def test_blob_round_trip_and_attach(fm):
    import torch
    rng = np.random.default_rng(5)
    text, alpha = _rand_text(rng, 70_000, 6, True)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=16))
    pats = _mixed_patterns(rng, text, alpha, 500, 16)
    d, o = fm.pack_patterns(pats)
    want_c = idx.count_batch(d, o)
    want_l = idx.locate_batch(d, o, limit=20)
    blob = idx.blob_to_host()
    idx2 = fm.FMIndex.from_host_blob(blob)
    assert (idx2.count_batch(d, o) == want_c).all()
    # replication path: device blob -> another device buffer (what the NCCL broadcast fills) -> attach
    ptr, nbytes = idx.blob()
    t = torch.empty(nbytes, dtype=torch.uint8, device="cuda:0")
    t.copy_(torch.from_numpy(blob))
    idx3 = fm.FMIndex.attach_blob(t.data_ptr(), nbytes, 0, keepalive=t)
    got = idx3.locate_batch(d, o, limit=20)
    for a, b in zip(got, want_l):
        assert (a == b).all()
    bad = blob.copy()
    bad[0] ^= 0xFF
    with pytest.raises(fm.CsfmError):
        fm.FMIndex.from_host_blob(bad)


def test_device_pointer_api(fm):
    import torch
    rng = np.random.default_rng(11)
    text, alpha = _rand_text(rng, 90_000, 4, True)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=8))
    orc = oracle.OracleIndex(text, stride=8)
    pats = _mixed_patterns(rng, text, alpha, 4000, 20)
    d, o = fm.pack_patterns(pats)
    npat = o.size - 1
    dev = torch.device("cuda:0")
    td = torch.from_numpy(d).to(dev)
    to = torch.from_numpy(o.astype(np.int64)).to(dev)
    tc = torch.zeros(npat, dtype=torch.int64, device=dev)
    tse = torch.zeros(2 * npat, dtype=torch.int64, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        idx.count_batch_device(td.data_ptr(), to.data_ptr(), npat, tc.data_ptr(), tse.data_ptr(), s.cuda_stream)
    s.synchronize()
    oc, ose = orc.count_batch(d, o)
    assert (tc.cpu().numpy().astype(np.uint64) == oc).all()
    assert (tse.cpu().numpy().astype(np.uint64).reshape(-1, 2) == ose).all()
    toffs = torch.zeros(npat + 1, dtype=torch.int64, device=dev)
    tst = torch.zeros(npat, dtype=torch.int32, device=dev)
    total = idx.locate_batch_device(td.data_ptr(), to.data_ptr(), npat, 9, toffs.data_ptr(), 0, 0, tst.data_ptr(), s.cuda_stream)
    tpos = torch.zeros(max(1, total), dtype=torch.int64, device=dev)
    total2 = idx.locate_batch_device(td.data_ptr(), to.data_ptr(), npat, 9, toffs.data_ptr(), tpos.data_ptr(), total,
                                     tst.data_ptr(), s.cuda_stream)
    s.synchronize()
    ooffs, opos, ostatus, _ = orc.locate_batch(d, o, limit=9)
    assert total == total2 == int(ooffs[-1])
    assert (toffs.cpu().numpy().astype(np.uint64) == ooffs).all()
    assert (tpos.cpu().numpy().astype(np.uint64)[:total] == opos).all()
    assert (tst.cpu().numpy() == ostatus).all()


def test_unaligned_and_odd_sized_device_batches(fm):
    """Pattern staging uses 16-byte aligned TMA bulk copies: unaligned batches, batches whose size
    is not a multiple of the 32-query chunk, empty-only chunks and oversized chunks must all take
    the direct path and give the same answers."""
    import torch
    rng = np.random.default_rng(21)
    text, alpha = _rand_text(rng, 120_000, 7, True)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=8))
    orc = oracle.OracleIndex(text, stride=8)
    dev = torch.device("cuda:0")
    for npat, maxlen, shift_b, shift_o in [(1, 9, 0, 0), (31, 9, 0, 0), (33, 9, 3, 0), (64, 9, 0, 1), (1000, 9, 5, 1),
                                           (4097, 40, 0, 0), (257, 300, 0, 0), (96, 0, 0, 0)]:
        pats = [text[s:s + m].tobytes() for s, m in zip(rng.integers(0, 100_000, npat), rng.integers(0, maxlen + 1, npat))]
        d, o = fm.pack_patterns(pats)
        oc, ose = orc.count_batch(d, o)
        # place the arrays at deliberately misaligned device addresses
        tb = torch.zeros(d.size + 64, dtype=torch.uint8, device=dev)
        tb[shift_b:shift_b + d.size] = torch.from_numpy(d).to(dev)
        to = torch.zeros(o.size + 8, dtype=torch.int64, device=dev)
        to[shift_o:shift_o + o.size] = torch.from_numpy(o.astype(np.int64)).to(dev)
        tc = torch.zeros(npat, dtype=torch.int64, device=dev)
        tse = torch.zeros(2 * npat, dtype=torch.int64, device=dev)
        idx.count_batch_device(tb.data_ptr() + shift_b, to.data_ptr() + 8 * shift_o, npat, tc.data_ptr(), tse.data_ptr(), 0)
        torch.cuda.synchronize()
        assert (tc.cpu().numpy().astype(np.uint64) == oc).all(), (npat, maxlen, shift_b, shift_o)
        assert (tse.cpu().numpy().astype(np.uint64).reshape(-1, 2) == ose).all()


def test_tma_staged_count_variant(fm, monkeypatch):
    """The TMA-staged count kernel (CSFM_PATTERN_STAGING=tma: cp.async.bulk + mbarrier double
    buffer per warp) must agree with the oracle exactly like the default direct-load kernel."""
    monkeypatch.setenv("CSFM_PATTERN_STAGING", "tma")
    rng = np.random.default_rng(77)
    for sigma, n, flags in [(4, 150_000, fm.BUILD_LAYOUT_NIBBLE128), (200, 200_000, 0)]:  # the staged variant is a layout-2 kernel
        text, alpha = _rand_text(rng, n, sigma, True)
        idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=16), flags=flags)
        assert idx.info().layout == 2
        orc = oracle.OracleIndex(text, stride=16)
        pats = _mixed_patterns(rng, text, alpha, 20_000, 40)
        pats += [text[s:s + m].tobytes() for s, m in zip(rng.integers(0, n - 3000, 40), rng.integers(65, 2500, 40))]
        d, o = fm.pack_patterns(pats)
        idx.set_instrumentation(1)
        counts, sp_ep = idx.count_batch(d, o, want_intervals=True)
        oc, ose, osteps = orc.count_batch(d, o, want_steps=True)
        assert (counts == oc).all() and (sp_ep == ose).all()
        assert idx.last_call_stats().search_steps == int(osteps.sum())
        offs, pos, status = idx.locate_batch(d, o, limit=11)
        ooffs, opos, ostatus, _ = orc.locate_batch(d, o, limit=11)
        assert (offs == ooffs).all() and (pos == opos).all() and (status == ostatus).all()


def test_async_submit_wait(fm):
    """csfm_count_batch_submit/_wait: several batches in flight on internal streams, pinned buffers."""
    import torch
    rng = np.random.default_rng(31)
    text, alpha = _rand_text(rng, 100_000, 4, True)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=16))
    orc = oracle.OracleIndex(text, stride=16)
    jobs = []
    for k in range(7):
        pats = _mixed_patterns(rng, text, alpha, 5000 + 37 * k, 18)
        d, o = fm.pack_patterns(pats)
        hb = torch.from_numpy(d.copy()).pin_memory() if d.size else torch.zeros(1, dtype=torch.uint8).pin_memory()
        ho = torch.from_numpy(o.astype(np.int64)).pin_memory()
        hc = torch.zeros(o.size - 1, dtype=torch.int64).pin_memory()
        hs = torch.zeros(2 * (o.size - 1), dtype=torch.int64).pin_memory()
        t = idx.count_batch_submit(hb.data_ptr(), ho.data_ptr(), o.size - 1, hc.data_ptr(), hs.data_ptr())
        jobs.append((t, d, o, hb, ho, hc, hs))
    for t, d, o, hb, ho, hc, hs in reversed(jobs):  # any order
        idx.count_batch_wait(t)
        idx.count_batch_wait(t)  # idempotent
        oc, ose = orc.count_batch(d, o)
        assert (hc.numpy().astype(np.uint64) == oc).all()
        assert (hs.numpy().astype(np.uint64).reshape(-1, 2) == ose).all()


@pytest.mark.parametrize("force_text", [False, True])
def test_async_submit32_compact(fm, force_text):
    """csfm_count_batch_submit32: u32 offsets in, u32 counts out, mixed with the u64 form on the same slots."""
    import torch
    rng = np.random.default_rng(77)
    text, alpha = _rand_text(rng, 150_000, 90, True)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=16), flags=fm.BUILD_FORCE_TEXT_CHECK if force_text else 0)
    orc = oracle.OracleIndex(text, stride=16)
    jobs = []
    for k in range(8):
        pats = _mixed_patterns(rng, text, alpha, 3000 + 101 * k, 24) if k != 5 else [b""]
        d, o = fm.pack_patterns(pats)
        hb = torch.from_numpy(d.copy()).pin_memory() if d.size else torch.zeros(1, dtype=torch.uint8).pin_memory()
        npat = o.size - 1
        if k % 3 == 2:
            ho = torch.from_numpy(o.astype(np.int64)).pin_memory()
            hc = torch.zeros(npat, dtype=torch.int64).pin_memory()
            t = idx.count_batch_submit(hb.data_ptr(), ho.data_ptr(), npat, hc.data_ptr())
        else:
            ho = torch.from_numpy(o.astype(np.uint32).view(np.int32)).pin_memory()
            hc = torch.full((npat,), -1, dtype=torch.int32).pin_memory()
            t = idx.count_batch_submit32(hb.data_ptr(), ho.data_ptr(), npat, hc.data_ptr())
        jobs.append((t, d, o, hb, ho, hc))
    for t, d, o, hb, ho, hc in jobs:
        idx.count_batch_wait(t)
        oc, _ = orc.count_batch(d, o)
        got = hc.numpy()
        got = got.view(np.uint32) if got.dtype == np.int32 else got
        assert (got.astype(np.uint64) == oc).all()


def test_replicate_handle(fm):
    """csfm_replicate: a second handle over its own copy of the blob (same device here; a peer device when the
    box has one) answers like the original, also after the original is destroyed."""
    import torch
    rng = np.random.default_rng(9)
    text, alpha = _rand_text(rng, 120_000, 70, True)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=4), flags=fm.BUILD_FORCE_TEXT_CHECK)
    orc = oracle.OracleIndex(text, stride=4)
    d, o = fm.pack_patterns(_mixed_patterns(rng, text, alpha, 5000, 24))
    oc, ose = orc.count_batch(d, o)
    ooffs, opos, ostatus, _ = orc.locate_batch(d, o, limit=9)
    reps = [idx.replicate(dev) for dev in range(min(2, torch.cuda.device_count()))] + [idx.replicate(0)]
    assert all(r.info().blob_bytes == idx.info().blob_bytes for r in reps)
    assert [r.info().device for r in reps][0] == 0 and reps[-1].blob()[0] != idx.blob()[0]
    idx.close()
    for r in reps:
        assert (r.count_batch(d, o) == oc).all()
        c, se = r.count_batch(d, o, want_intervals=True)
        assert (c == oc).all() and (se == ose).all()
        offs, pos, status = r.locate_batch(d, o, limit=9)
        assert (offs == ooffs).all() and (pos == opos).all() and (status == ostatus).all()
        r.close()


def test_concurrent_host_threads_one_handle(fm):
    """SURVEY §8b threading row: queries on ONE handle from several host threads (ctypes releases the GIL,
    so the calls really overlap): host-pointer count and locate, device-pointer count on private streams and
    the streaming form, all against the same oracle answers."""
    import threading
    import torch
    rng = np.random.default_rng(5)
    text, alpha = _rand_text(rng, 200_000, 40, True)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=8), flags=fm.BUILD_FORCE_TEXT_CHECK)
    orc = oracle.OracleIndex(text, stride=8)
    work = []
    for k in range(6):
        d, o = fm.pack_patterns(_mixed_patterns(rng, text, alpha, 4000 + 50 * k, 20))
        oc, ose = orc.count_batch(d, o)
        ooffs, opos, ostatus, _ = orc.locate_batch(d, o, limit=7)
        work.append((d, o, oc, ose, ooffs, opos, ostatus))
    errors = []

    def host_worker(k):
        try:
            d, o, oc, ose, ooffs, opos, ostatus = work[k]
            for _ in range(6):
                assert (idx.count_batch(d, o) == oc).all()
                c, se = idx.count_batch(d, o, want_intervals=True)
                assert (c == oc).all() and (se == ose).all()
                offs, pos, status = idx.locate_batch(d, o, limit=7)
                assert (offs == ooffs).all() and (pos == opos).all() and (status == ostatus).all()
        except BaseException as e:  # noqa: BLE001 - reported by the main thread
            errors.append(repr(e))

    def device_worker(k):
        try:
            d, o, oc = work[k][:3]
            torch.cuda.set_device(0)
            stream = torch.cuda.Stream()
            db = torch.from_numpy(d).cuda()
            do = torch.from_numpy(o.astype(np.int64)).cuda()
            dc = torch.zeros(o.size - 1, dtype=torch.int64, device="cuda")
            for _ in range(20):
                dc.zero_()
                stream.wait_stream(torch.cuda.current_stream())
                idx.count_batch_device(db.data_ptr(), do.data_ptr(), o.size - 1, dc.data_ptr(), 0, stream.cuda_stream)
                stream.synchronize()
                assert (dc.cpu().numpy().astype(np.uint64) == oc).all()
        except BaseException as e:  # noqa: BLE001
            errors.append(repr(e))

    def stream_worker(k):
        try:
            d, o, oc = work[k][:3]
            hb = torch.from_numpy(d.copy()).pin_memory()
            hl = torch.from_numpy(np.diff(o).astype(np.uint8)).pin_memory()
            hc = torch.zeros(o.size - 1, dtype=torch.int32).pin_memory()
            for _ in range(10):
                hc.fill_(-1)
                idx.count_batch_wait(idx.count_batch_submit_len8(hb.data_ptr(), int(o[-1]), hl.data_ptr(), o.size - 1, hc.data_ptr()))
                assert (hc.numpy().view(np.uint32).astype(np.uint64) == oc).all()
        except BaseException as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=host_worker, args=(0,)), threading.Thread(target=host_worker, args=(1,)),
               threading.Thread(target=device_worker, args=(2,)), threading.Thread(target=device_worker, args=(3,)),
               threading.Thread(target=stream_worker, args=(4,)), threading.Thread(target=stream_worker, args=(5,))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("force_text", [False, True])
def test_async_submit_len8(fm, force_text):
    """csfm_count_batch_submit_len8: one length byte per pattern in (offsets rebuilt by a device prefix sum),
    u32 counts out; empty patterns, 255-byte patterns, an empty batch and a wrong nbytes."""
    import torch
    rng = np.random.default_rng(78)
    text, alpha = _rand_text(rng, 120_000, 60, True)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=16), flags=fm.BUILD_FORCE_TEXT_CHECK if force_text else 0)
    orc = oracle.OracleIndex(text, stride=16)
    jobs = []
    for k in range(6):
        pats = _mixed_patterns(rng, text, alpha, 2500 + 77 * k, 24)
        pats += [b"", text[1000:1255].tobytes(), text[-255:].tobytes(), b""]
        if k == 4:
            pats = [b""]
        d, o = fm.pack_patterns(pats)
        npat = o.size - 1
        hb = torch.from_numpy(d.copy()).pin_memory() if d.size else torch.zeros(1, dtype=torch.uint8).pin_memory()
        hl = torch.from_numpy(np.diff(o).astype(np.uint8)).pin_memory()
        hc = torch.full((npat,), -1, dtype=torch.int32).pin_memory()
        t = idx.count_batch_submit_len8(hb.data_ptr(), int(o[-1]), hl.data_ptr(), npat, hc.data_ptr())
        jobs.append((t, d, o, hb, hl, hc))
    for t, d, o, hb, hl, hc in jobs:
        idx.count_batch_wait(t)
        oc, _ = orc.count_batch(d, o)
        assert (hc.numpy().view(np.uint32).astype(np.uint64) == oc).all()
    t, d, o, hb, hl, hc = jobs[0]
    with pytest.raises(fm.CsfmError):
        idx.count_batch_submit_len8(hb.data_ptr(), int(o[-1]) + 1, hl.data_ptr(), o.size - 1, hc.data_ptr())
    idx.count_batch_wait(idx.count_batch_submit_len8(0, 0, 0, 0, 0))  # empty batch: a ticket, no work


@pytest.mark.parametrize("sigma,n,term", [(4, 300_000, True), (255, 400_000, True), (20, 50_000, True), (4, 100_000, False)])
def test_large_table_build(fm, sigma, n, term):
    """CSFM_BUILD_LARGE_TABLE: a long k-mer key, then straight to the text verification (no rank step when
    the lookup already leaves at most four rows). Same counts, intervals and positions as the oracle;
    without a unique terminator the flag only enlarges the table."""
    rng = np.random.default_rng(sigma * 11 + n)
    text, alpha = _rand_text(rng, n, sigma, term)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=8), flags=fm.BUILD_LARGE_TABLE)
    info = idx.info()
    assert info.text_check == (1 if term else 0)
    small = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=8))
    assert info.kmer_k >= small.info().kmer_k
    orc = oracle.OracleIndex(text, stride=8)
    pats = _mixed_patterns(rng, text, alpha, 30_000, 40)
    pats += [text[s:s + m].tobytes() for m in range(1, 12) for s in (0, 1, n // 2, n - m)]
    d, o = fm.pack_patterns(pats)
    idx.set_instrumentation(1)
    counts = idx.count_batch(d, o)
    if term:
        assert idx.last_call_stats().text_checks > 0
    idx.set_instrumentation(0)
    oc, ose = orc.count_batch(d, o)
    assert (counts == oc).all()
    c2, se = idx.count_batch(d, o, want_intervals=True)
    assert (c2 == oc).all() and (se == ose).all()
    offs, pos, status = idx.locate_batch(d, o, limit=20)
    ooffs, opos, ostatus, _ = orc.locate_batch(d, o, limit=20)
    assert (offs == ooffs).all() and (status == ostatus).all() and (pos == opos).all()


@pytest.mark.parametrize("sigma,n", [(4, 300_000), (255, 400_000), (20, 50_000)])
def test_text_verification_shortcut(fm, sigma, n):
    """Counts WITHOUT intervals take the shortcut: once a query's interval is a single row, its
    remaining characters are compared with the text. Same counts as the oracle and as an index
    built without the text sections, for hits, near misses, cyclic wrap-arounds, patterns with the
    terminator inside, long patterns and bytes outside the alphabet."""
    rng = np.random.default_rng(sigma * 7 + n)
    alpha = np.sort(rng.choice(np.arange(1, 256), sigma, replace=False)).astype(np.uint8)
    text = np.concatenate([alpha[rng.integers(0, sigma, n - 1)], np.zeros(1, np.uint8)]).astype(np.uint8)
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=32), flags=fm.BUILD_FORCE_TEXT_CHECK)
    assert idx.info().text_check == 1
    assert fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=32)).info().text_check == 0  # small: levels fit in L2
    plain = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=32), flags=fm.BUILD_NO_TEXT_CHECK)
    assert plain.info().text_check == 0
    orc = oracle.OracleIndex(text, stride=32)
    pats = []
    for _ in range(6000):
        m = int(rng.integers(1, 41))
        s = int(rng.integers(0, n - m))
        p = text[s:s + m].copy()
        r = rng.random()
        if r < 0.35 and m > 1:
            p[int(rng.integers(0, m))] = alpha[rng.integers(0, sigma)]      # near miss anywhere
        elif r < 0.45:
            p[0] = 0xFF                                                    # byte outside the alphabet, at the far end
        pats.append(p.tobytes())
    for k in (1, 2, 5, 17, 31, 40):                                        # cyclic: ... end of text, terminator, start of text
        for t in (1, 3, 9, 30):
            pats.append(np.concatenate([text[n - k:], text[:t]]).tobytes())
    pats += [text[s:s + m].tobytes() for s, m in zip(rng.integers(0, n - 3000, 30), rng.integers(60, 2500, 30))]
    pats += [text[:m].tobytes() for m in (1, 2, 8, 33, 70)] + [b"", text[-1:].tobytes()]
    d, o = fm.pack_patterns(pats)
    oc, _ = orc.count_batch(d, o)
    idx.set_instrumentation(1)
    got = idx.count_batch(d, o)
    st = idx.last_call_stats()
    assert (got == oc).all()
    assert st.text_checks > 1000                       # the shortcut really ran
    assert st.kernel_launches == 1
    idx.set_instrumentation(0)
    assert (idx.count_batch(d, o) == oc).all()         # the uninstrumented kernel
    assert (plain.count_batch(d, o) == oc).all()
    # the two-pass form (one query per thread, then the sub-warp kernel over what it left; opt-in, measured
    # slower on B200: profiles/README.md) must agree
    os.environ["CSFM_TWO_PASS"] = "1"
    try:
        two = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=32), flags=fm.BUILD_FORCE_TEXT_CHECK)
    finally:
        del os.environ["CSFM_TWO_PASS"]
    two.set_instrumentation(1)
    assert (two.count_batch(d, o) == oc).all()
    assert two.last_call_stats().kernel_launches == 2 and two.last_call_stats().text_checks > 1000
    two.set_instrumentation(0)
    assert (two.count_batch(d, o) == oc).all()
    # a device batch that is not 4-byte aligned, and a pattern that ends on the last byte of the batch
    import torch
    dev = torch.device("cuda", 0)
    for shift in (1, 2, 3, 5):
        tb = torch.zeros(d.size + 8, dtype=torch.uint8, device=dev)
        tb[shift: shift + d.size] = torch.from_numpy(d).to(dev)
        to = torch.from_numpy(o.astype(np.int64)).to(dev)
        for ix in (idx, two):
            tc = torch.zeros(len(pats), dtype=torch.int64, device=dev)
            ix.count_batch_device(tb.data_ptr() + shift, to.data_ptr(), len(pats), tc.data_ptr(), 0, 0)
            torch.cuda.synchronize()
            assert (tc.cpu().numpy().astype(np.uint64) == oc).all(), shift
    # with intervals requested the shortcut is off and the intervals are exact
    c2, se2 = idx.count_batch(d, o, want_intervals=True)
    oc2, ose2 = orc.count_batch(d, o)
    assert (c2 == oc2).all() and (se2 == ose2).all()
    # locate on an index that carries its suffix array: positions by one gather per row (no LF walk),
    # same SA-row order, same limit semantics, same positions as the walking index and the oracle
    for limit in (3, 100000):
        offs, pos, status = idx.locate_batch(d, o, limit=limit)
        assert idx.last_call_stats().lf_steps == 0
        ooffs, opos, ostatus, _ = orc.locate_batch(d, o, limit=limit)
        assert (offs == ooffs).all() and (pos == opos).all() and (status == ostatus).all()
        poffs, ppos, pstatus = plain.locate_batch(d, o, limit=limit)
        assert (poffs == ooffs).all() and (ppos == opos).all() and (pstatus == ostatus).all()


@pytest.mark.parametrize("copies,sigma", [(2, 200), (3, 60), (4, 30), (7, 120)])
def test_text_verification_of_several_rows(fm, copies, sigma):
    """Near-repeats: the text is `copies` mutated copies of one block, so most patterns narrow to 2..copies
    rows that the remaining characters then tell apart (up to four rows are verified lane by lane,
    more keep stepping). Counts 0..copies all occur."""
    rng = np.random.default_rng(copies * 1000 + sigma)
    block = rng.integers(1, sigma + 1, 20_000).astype(np.uint8)
    parts = []
    for _ in range(copies):
        b = block.copy()
        hit = rng.random(b.size) < 0.02
        b[hit] = rng.integers(1, sigma + 1, int(hit.sum())).astype(np.uint8)
        parts.append(b)
    text = np.concatenate(parts + [np.zeros(1, np.uint8)]).astype(np.uint8)
    n = text.size
    idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=16), flags=fm.BUILD_FORCE_TEXT_CHECK)
    assert idx.info().text_check == 1
    orc = oracle.OracleIndex(text, stride=16)
    pats = []
    for _ in range(8000):
        m = int(rng.integers(4, 45))
        s = int(rng.integers(0, n - m))
        p = text[s:s + m].copy()
        if rng.random() < 0.2:
            p[int(rng.integers(0, m))] = rng.integers(1, sigma + 1)
        pats.append(p.tobytes())
    d, o = fm.pack_patterns(pats)
    oc, _ = orc.count_batch(d, o)
    idx.set_instrumentation(1)
    got = idx.count_batch(d, o)
    assert (got == oc).all()
    assert idx.last_call_stats().text_checks > 1000
    seen = set(np.unique(oc).tolist())
    assert set(range(0, min(copies, 4) + 1)) <= seen


def test_no_text_check_without_a_unique_smallest_terminator(fm):
    for text in (b"banana", b"abab$abab$", b"zzz\x01zzz\x00\x00"):
        idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=2), flags=fm.BUILD_FORCE_TEXT_CHECK)
        assert idx.info().text_check == 0
    assert fm.FMIndex.build_from_text(b"banana$", fm.BuildParams(ssa_stride=2), flags=fm.BUILD_FORCE_TEXT_CHECK).info().text_check == 1
