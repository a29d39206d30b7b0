"""CPU: the bookkeeping of bench.py that decides which evidence a bench line may carry: a committed ncu capture is
used only when it was taken from the kernel sources in this tree AND names the kernel the run launches (a stale file must
not decorate a new kernel), and the L2 / working-set ceilings come from the committed probe."""
import json
import os

import pytest

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_capture_matches_these_sources():
    doc = json.load(open(os.path.join(ROOT, "profiles", "count_kernel_traffic.json")))
    if doc["source_sha16"] != bench.kernel_sources_sha16():
        pytest.skip("csrc/ changed since the captures were taken: bench.py reports traffic = null until "
                    "tools/profile_round.sh + tools/profile_collect.sh are re-run on a B200")
    for key, kernel in (("c3", "count2_kernel<true,false>"), ("c3_stepping", "count2_kernel<false,false>"), ("c2", "count3_kernel"),
                        ("c5", "count3_kernel"), ("c4_walk", "walk3_kernel")):
        bytes_, meta = bench.committed_traffic(key, kernel)
        assert meta["valid"] and bytes_ > 0, (key, meta)


def test_stale_or_foreign_captures_are_dropped(tmp_path, monkeypatch):
    doc = json.load(open(os.path.join(ROOT, "profiles", "count_kernel_traffic.json")))
    prof = tmp_path / "profiles"
    prof.mkdir()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    monkeypatch.setattr(bench, "kernel_sources_sha16", lambda: doc["source_sha16"])
    (prof / "count_kernel_traffic.json").write_text(json.dumps(doc))
    assert bench.committed_traffic("c3", "count2_kernel<true,false>")[0] > 0
    # a capture of another kernel under the same key
    assert bench.committed_traffic("c3", "count3_kernel")[0] is None
    # a key that was never captured
    got, meta = bench.committed_traffic("c9", "count2_kernel")
    assert got is None and not meta["valid"]
    # the sources moved on: every capture is stale
    stale = dict(doc, source_sha16="0" * 16)
    (prof / "count_kernel_traffic.json").write_text(json.dumps(stale))
    got, meta = bench.committed_traffic("c3", "count2_kernel<true,false>")
    assert got is None and "belongs to sources" in meta["why"]
    # no file at all
    os.remove(prof / "count_kernel_traffic.json")
    assert bench.committed_traffic("c3", "count2_kernel<true,false>")[0] is None


def test_working_set_ceilings_come_from_the_probe():
    small, src = bench.l2_random_peak(30 << 20, 64)
    big, _ = bench.l2_random_peak(1 << 30, 64)
    mid, _ = bench.l2_random_peak(123 << 20, 64)
    assert "r2_l2_sweep_probe.json" in src
    assert small > mid > big > 0          # the fetch rate falls as the working set outgrows the L2
    assert bench.l2_random_peak(30 << 20, 128)[0] > bench.l2_random_peak(300 << 20, 128)[0]
    # between two rows of the sweep the ceiling is interpolated, not taken from the next larger row; beyond the sweep the
    # last row stands (out of HBM the rate no longer depends on the size)
    lo, _ = bench.l2_random_peak(160 << 20, 64)
    hi, _ = bench.l2_random_peak(192 << 20, 64)
    between, src2 = bench.l2_random_peak(176 << 20, 64)
    assert lo > between > hi and "interpolated" in src2
    assert abs(between - (lo + hi) / 2) < 1e-6 * lo
    beyond, src3 = bench.l2_random_peak(4 << 30, 64)
    assert beyond == big and "largest" in src3


def test_count_working_set_is_lines_plus_table():
    from types import SimpleNamespace as NS
    # C2 in the marked form: 2^26 / 128 + 1 lines of 64 bytes, 4^11 entries of 8 bytes; the samples are not part of it
    c2 = NS(blocks_per_level=(1 << 26) // 128 + 1, line_bytes=64, levels=1, kmer_k=11, sigma=5, layout=3, text_check=0, half_table=0)
    assert bench.count_working_set(c2) == ((1 << 26) // 128 + 1) * 64 + 4 ** 11 * 8
    # C3: two levels of 128-byte lines, 256^3 keys of 4 bytes (tiled table) and the half-step table
    c3 = NS(blocks_per_level=(1 << 30) // 128 + 1, line_bytes=128, levels=2, kmer_k=3, sigma=256, layout=2, text_check=1, half_table=1)
    assert bench.count_working_set(c3) == ((1 << 30) // 128 + 1) * 256 + 256 ** 3 * (4 + 128)
