"""GPU, BASELINE.json configs[2] and configs[3] at FULL size. No CPU index of that size exists (the oracle's
restatement at n = 2^30 would take minutes and tens of GB), so the expected values come from an INDEPENDENT checker,
tests/sa_checker.py: the GPU-built suffix array is certified by the O(n) checker (a certified SA is THE reference SA:
suffixes are distinct and sais.hpp:13's order is total), and for a text with a unique smallest terminator the reference's
interval of a pattern is [lower_bound, upper_bound) among the suffixes and locate reports SA[sp], SA[sp+1], ... — both
bounds by binary search with direct text comparisons in plain torch; nothing of the engine's index takes part.

  C3  n = 2^30 bytes, sigma = 256, 1 M text-sampled patterns of length 8..32 (the bench workload)
      * SA certificate; BWT, SSA and C equal their definitions over it (bwt.hpp:10-13, fm_index.cpp:36-65);
      * all 10^6 counts of the default kernel (tables + text verification) and of the stepping kernel, and all 10^6
        intervals, equal the checker's; a second batch with misses, 1..3-byte patterns and a 200-byte pattern too;
      * located positions equal SA[sp .. sp + min(count, limit)) IN ORDER (limit 100000 and 2);
      * (kept from round 1) default kernel == stepping kernel, counts and sorted positions == a naive scan of the text.
  C4  n = 2^28 DNA + '$', ssa_stride 32, the config's 1 M text-sampled patterns of length 10 (~257 occurrences each)
      * SA certificate; all counts and intervals == the checker's; all 2.6e8 located positions IN ORDER;
      * the length-12 set with limit 7; `limit` keeps the FIRST rows; naive-scan spot checks;
      * the LF-walking index and the index with a resident suffix array report identical positions.
"""
import pytest

pytestmark = pytest.mark.gpu


def _naive_positions(text, pat):
    """All start positions of `pat` (1-D uint8 tensor) in `text`, by direct comparison on the device."""
    import torch
    n, m = text.numel(), pat.numel()
    hit = text[: n - m + 1] == pat[0]
    for k in range(1, m):
        hit &= text[k: n - m + 1 + k] == pat[k]
    return torch.nonzero(hit).flatten()


def _locate_device(idx, bytes_d, offs_d, npat, limit, dev):
    import torch
    offs = torch.zeros(npat + 1, dtype=torch.int64, device=dev)
    status = torch.zeros(npat, dtype=torch.int32, device=dev)
    tot = idx.locate_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, limit, offs.data_ptr(), 0, 0, status.data_ptr())
    pos = torch.zeros(max(1, tot), dtype=torch.int64, device=dev)
    idx.locate_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, limit, offs.data_ptr(), pos.data_ptr(), tot,
                            status.data_ptr())
    torch.cuda.synchronize()
    return offs, pos[:tot], status


def _positions_are_occurrences(text, bytes_d, offs_d, offs, pos, max_len):
    import torch
    tot = pos.numel()
    q_of = torch.searchsorted(offs, torch.arange(tot, device=pos.device), right=True) - 1
    lens = (offs_d[1:] - offs_d[:-1])[q_of]
    ok = torch.ones(tot, dtype=torch.bool, device=pos.device)
    for k in range(max_len):
        live = lens > k
        t = text[torch.where(live, pos + k, torch.zeros_like(pos))]
        p = bytes_d[torch.where(live, offs_d[q_of] + k, torch.zeros_like(pos))]
        ok &= (~live) | (t == p)
    return bool(ok.all().item())


@pytest.fixture(scope="module")
def c3():
    import torch
    import csfm_b200 as fm
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n = 1 << 30
    text = fm.workloads.byte_text_torch(n, 3, dev)
    idx = fm.FMIndex.build_from_text_device(text.data_ptr(), n, fm.BuildParams(ssa_stride=32), device=0, flags=fm.BUILD_KEEP_SA)
    yield fm, dev, text, idx
    idx.close()
    del text
    torch.cuda.empty_cache()


def test_c3_build_products_certified(c3):
    import numpy as np
    import torch
    fm, dev, text, idx = c3
    n = text.numel()
    info = idx.info()
    assert info.levels == 2 and info.line_bytes == 128 and info.sigma == 256 and info.text_check == 1
    cert = fm.workloads.certify_sa_torch(text, idx.sa_device_ptr(), n)
    assert cert["ok"], cert

    class _Mem:
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    sa = torch.as_tensor(_Mem(idx.sa_device_ptr(), 4 * n), device=dev).view(torch.int32)
    # SSA: samples[k] = SA[k * stride]  (fm_index.cpp:57-65)
    assert (idx.ssa() == sa[::32].contiguous().cpu().numpy().view(np.uint32)).all()
    # BWT[i] = T[(SA[i] - 1) mod n]  (bwt.hpp:10-13), compared in 2^27-row pieces
    bwt = torch.from_numpy(idx.bwt())
    for lo in range(0, n, 1 << 27):
        s = sa[lo: lo + (1 << 27)].to(torch.int64) & 0xFFFFFFFF
        want = text[(s - 1) % n]
        assert torch.equal(want.cpu(), bwt[lo: lo + (1 << 27)])
    # C[c] = number of symbols smaller than c  (fm_index.cpp:36-47)
    hist = torch.zeros(256, dtype=torch.int64, device=dev)
    for lo in range(0, n, 1 << 27):
        hist += torch.bincount(text[lo: lo + (1 << 27)].int(), minlength=256)
    C = np.zeros(257, np.uint32)
    C[1:] = np.cumsum(hist.cpu().numpy()).astype(np.uint32)
    assert (idx.C_array() == C).all()


def test_c3_counts_intervals_and_ordered_locate_vs_sa_checker(c3):
    """All 10^6 counts AND intervals, and located positions IN ORDER, against the independent checker
    (binary search over the certified suffix array with direct text comparisons, tests/sa_checker.py):
    nothing of the engine's index takes part in the expected values."""
    import torch
    from sa_checker import DeviceArrayU32, expected_locate, sa_intervals
    fm, dev, text, idx = c3
    n = text.numel()
    npat = 1_000_000
    bytes_d, offs_d = fm.workloads.sampled_patterns_torch(text, npat, 8, 32, 4, 5)
    sa = DeviceArrayU32.from_ptr(idx.sa_device_ptr(), n, dev)  # certified by test_c3_build_products_certified
    lb, ub = sa_intervals(text, sa, bytes_d, offs_d)
    fast = torch.zeros(npat, dtype=torch.int64, device=dev)
    plain = torch.zeros(npat, dtype=torch.int64, device=dev)
    spep = torch.zeros(2 * npat, dtype=torch.int64, device=dev)
    idx.count_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, fast.data_ptr(), 0)
    idx.count_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, plain.data_ptr(), spep.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(fast, ub - lb)     # default kernel (tables + text verification)
    assert torch.equal(plain, ub - lb)    # stepping kernel
    se = spep.view(-1, 2)
    assert torch.equal(se[:, 0], lb) and torch.equal(se[:, 1], ub)
    # a batch with misses and short patterns: random strings, 1..3-byte patterns, a pattern longer than 32
    from csfm_b200 import workloads as w
    import numpy as np
    rd, ro = w.random_patterns_np(bytes(range(1, 256)), 50_000, 5, 77)
    extra = [text[5:6], text[100:102], text[1000:1003], text[12345:12345 + 200], text[n - 9:n]]
    xb = torch.cat([torch.from_numpy(rd).to(dev)] + extra)
    xo = torch.cat([torch.from_numpy(ro.astype(np.int64)).to(dev),
                    int(ro[-1]) + torch.cumsum(torch.tensor([e.numel() for e in extra], device=dev), 0)])
    nx = xo.numel() - 1
    xlb, xub = sa_intervals(text, sa, xb, xo)
    xc = torch.zeros(nx, dtype=torch.int64, device=dev)
    xse = torch.zeros(2 * nx, dtype=torch.int64, device=dev)
    idx.count_batch_device(xb.data_ptr(), xo.data_ptr(), nx, xc.data_ptr(), 0)
    torch.cuda.synchronize()
    assert torch.equal(xc, xub - xlb)
    idx.count_batch_device(xb.data_ptr(), xo.data_ptr(), nx, xc.data_ptr(), xse.data_ptr())
    torch.cuda.synchronize()
    hit = xub > xlb
    assert torch.equal(xc, xub - xlb)
    assert torch.equal(xse.view(-1, 2)[hit, 0], xlb[hit]) and torch.equal(xse.view(-1, 2)[hit, 1], xub[hit])
    assert int(xse.view(-1, 2)[~hit].abs().sum()) == 0  # empty results are normalised to (0, 0)
    # locate: positions in SA-row order (sais.hpp:13 order, fm_index.cpp:125-153), limit honoured
    for limit in (100000, 2):
        q = 200_000
        offs, pos, status = _locate_device(idx, bytes_d, offs_d, q, limit, dev)
        e_offs, e_pos = expected_locate(sa, lb[:q], ub[:q], limit)
        assert int(status.max()) == 0
        assert torch.equal(offs, e_offs) and torch.equal(pos, e_pos)
    offs, pos, status = _locate_device(idx, xb, xo, nx, 1000, dev)
    e_offs, e_pos = expected_locate(sa, xlb, xub, 1000)
    assert int(status.max()) == 0 and torch.equal(offs, e_offs) and torch.equal(pos, e_pos)


def test_c3_one_million_counts(c3):
    import torch
    fm, dev, text, idx = c3
    npat = 1_000_000
    bytes_d, offs_d = fm.workloads.sampled_patterns_torch(text, npat, 8, 32, 4, 5)
    fast = torch.zeros(npat, dtype=torch.int64, device=dev)
    plain = torch.zeros(npat, dtype=torch.int64, device=dev)
    spep = torch.zeros(2 * npat, dtype=torch.int64, device=dev)
    idx.set_instrumentation(1)
    idx.count_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, fast.data_ptr(), 0)
    torch.cuda.synchronize()
    st = idx.last_call_stats()
    assert st.text_checks > 0.9 * npat           # the verification kernel really ran
    idx.count_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, plain.data_ptr(), spep.data_ptr())
    torch.cuda.synchronize()
    assert idx.last_call_stats().text_checks == 0  # asking for the intervals selects the stepping kernel
    idx.set_instrumentation(0)
    assert torch.equal(fast, plain)
    assert bool((fast >= 1).all())
    se = spep.view(-1, 2)
    assert torch.equal(se[:, 1] - se[:, 0], plain)   # the interval of a non-empty result has `count` rows
    # count and locate against a naive scan of the text
    offs, pos, status = _locate_device(idx, bytes_d, offs_d, 4096, 100000, dev)
    assert int(status.max()) == 0
    assert torch.equal(offs[1:] - offs[:-1], fast[:4096])
    assert _positions_are_occurrences(text, bytes_d, offs_d, offs, pos, 32)
    for q in range(0, 24):
        pat = bytes_d[int(offs_d[q]): int(offs_d[q + 1])]
        want = _naive_positions(text, pat)
        assert int(fast[q]) == want.numel()
        got = pos[int(offs[q]): int(offs[q + 1])]
        assert torch.equal(torch.sort(got).values, want)
    # short patterns (length 1..3) have many occurrences and stay in the stepping path
    short = [text[5:6], text[100:102], text[1000:1003]]
    sb = torch.cat(short)
    so = torch.tensor([0, 1, 3, 6], dtype=torch.int64, device=dev)
    sc = torch.zeros(3, dtype=torch.int64, device=dev)
    idx.count_batch_device(sb.data_ptr(), so.data_ptr(), 3, sc.data_ptr(), 0)
    torch.cuda.synchronize()
    for k, pat in enumerate(short):
        assert int(sc[k]) == _naive_positions(text, pat).numel()


@pytest.fixture(scope="module")
def c4():
    import torch
    import csfm_b200 as fm
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n = 1 << 28
    text = fm.workloads.dna_text_torch(n, 6, dev)
    idx = fm.FMIndex.build_from_text_device(text.data_ptr(), n, fm.BuildParams(ssa_stride=32), device=0, flags=fm.BUILD_KEEP_SA)
    yield fm, dev, text, idx
    idx.close()
    del text
    torch.cuda.empty_cache()


def test_c4_one_million_patterns_ordered_locate_vs_sa_checker(c4):
    """configs[3] at its stated size: 1 M text-sampled patterns of length 10 (~257 occurrences each, ~2.6e8
    positions), every count, interval and position IN ORDER against the independent checker."""
    import torch
    from sa_checker import DeviceArrayU32, expected_locate, sa_intervals
    fm, dev, text, idx = c4
    n = text.numel()
    cert = fm.workloads.certify_sa_torch(text, idx.sa_device_ptr(), n)
    assert cert["ok"], cert
    sa = DeviceArrayU32.from_ptr(idx.sa_device_ptr(), n, dev)
    npat, plen = 1_000_000, 10
    bytes_d, offs_d = fm.workloads.sampled_patterns_torch(text, npat, plen, plen, 0, 8)
    lb, ub = sa_intervals(text, sa, bytes_d, offs_d)
    counts = torch.zeros(npat, dtype=torch.int64, device=dev)
    spep = torch.zeros(2 * npat, dtype=torch.int64, device=dev)
    idx.count_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, counts.data_ptr(), spep.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(counts, ub - lb)
    assert torch.equal(spep.view(-1, 2)[:, 0], lb) and torch.equal(spep.view(-1, 2)[:, 1], ub)
    del spep
    offs, pos, status = _locate_device(idx, bytes_d, offs_d, npat, 100000, dev)
    assert int(status.max()) == 0
    e_offs, e_pos = expected_locate(sa, lb, ub, 100000)
    assert torch.equal(offs, e_offs)
    assert torch.equal(pos, e_pos)
    del pos, e_pos
    # secondary set of the config: length 12 (~17 occurrences each), small limit
    b12, o12 = fm.workloads.sampled_patterns_torch(text, 200_000, 12, 12, 0, 9)
    lb12, ub12 = sa_intervals(text, sa, b12, o12)
    offs, pos, status = _locate_device(idx, b12, o12, 200_000, 7, dev)
    e_offs, e_pos = expected_locate(sa, lb12, ub12, 7)
    assert int(status.max()) == 0 and torch.equal(offs, e_offs) and torch.equal(pos, e_pos)


def test_c4_locate_full_size(c4):
    import torch
    fm, dev, text, idx = c4
    n = text.numel()
    info = idx.info()
    assert info.levels == 1 and info.ssa_stride == 32 and info.nsamp == n // 32
    npat, plen = 50_000, 10
    bytes_d, offs_d = fm.workloads.sampled_patterns_torch(text, npat, plen, plen, 0, 8)
    counts = torch.zeros(npat, dtype=torch.int64, device=dev)
    idx.count_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, counts.data_ptr(), 0)
    offs, pos, status = _locate_device(idx, bytes_d, offs_d, npat, 100000, dev)
    assert int(status.max()) == 0
    assert int(counts.max()) < 100000 and int(counts.sum()) == pos.numel()
    assert torch.equal(offs[1:] - offs[:-1], counts)
    assert _positions_are_occurrences(text, bytes_d, offs_d, offs, pos, plen)
    for q in range(16):
        want = _naive_positions(text, bytes_d[q * plen: (q + 1) * plen])
        assert int(counts[q]) == want.numel()
        assert torch.equal(torch.sort(pos[int(offs[q]): int(offs[q + 1])]).values, want)
    # `limit` keeps the FIRST rows of the interval (fm_index.cpp:118-120)
    offs5, pos5, _ = _locate_device(idx, bytes_d, offs_d, 1000, 5, dev)
    for q in range(0, 1000, 97):
        k = min(5, int(counts[q]))
        assert int(offs5[q + 1] - offs5[q]) == k
        assert torch.equal(pos5[int(offs5[q]): int(offs5[q + 1])], pos[int(offs[q]): int(offs[q]) + k])
    # the index that carries its whole suffix array reads SA[row]: same positions, same order
    idx_sa = fm.FMIndex.build_from_text_device(text.data_ptr(), n, fm.BuildParams(ssa_stride=32), device=0,
                                               flags=fm.BUILD_FORCE_TEXT_CHECK)
    assert idx_sa.info().text_check == 1
    offs2, pos2, status2 = _locate_device(idx_sa, bytes_d, offs_d, npat, 100000, dev)
    assert torch.equal(offs2, offs) and torch.equal(pos2, pos) and int(status2.max()) == 0
    idx_sa.close()
    # the default index samples the suffix array by text position (marked lines, walks of at most stride - 1 steps); the
    # reference's row-sampled walk (CSFM_BUILD_ROW_SAMPLES) gives the same positions in the same order
    assert info.position_samples == 1 and info.blocks_per_level == n // 128 + 1
    idx.set_instrumentation(1)
    _locate_device(idx, bytes_d, offs_d, npat, 100000, dev)
    assert idx.last_call_stats().lf_steps == int((pos % 32).sum())
    idx.set_instrumentation(0)
    rows = fm.FMIndex.build_from_text_device(text.data_ptr(), n, fm.BuildParams(ssa_stride=32), device=0, flags=fm.BUILD_ROW_SAMPLES)
    assert rows.info().position_samples == 0 and rows.info().blocks_per_level == n // 192 + 1
    offs3, pos3, status3 = _locate_device(rows, bytes_d, offs_d, npat, 100000, dev)
    assert torch.equal(offs3, offs) and torch.equal(pos3, pos) and int(status3.max()) == 0
    rows.close()
