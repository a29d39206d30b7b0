"""GPU, BASELINE.json configs[4] at FULL size: a 4e9-byte DNA text + '$', suffix array / BWT / index built on
the GPU (cs::FMIndex::build_from_text, fm_index.cpp:16-69; order of build_sa_naive, sais.hpp:8-16), then

  * the suffix array passes the O(n) certificate (a certified SA is THE reference SA);
  * BWT and SSA samples equal their definitions over it (bwt.hpp:10-13, fm_index.cpp:57-65), on sampled rows;
  * 10^6 text-sampled patterns of length 20 (the config's sweep), 2*10^5 random strings (mostly misses) and
    frequent patterns of length 12: every count AND interval equals the independent checker
    (tests/sa_checker.py: binary search over the certified SA with direct text comparisons);
  * located positions equal SA[sp], SA[sp+1], ... IN ORDER, with and without a small limit
    (fm_index.cpp:107-157).

Needs one GPU with more than 150 GB free (the prefix-doubling sort holds 33 n bytes of temporaries); skipped
otherwise.
"""
import pytest

pytestmark = pytest.mark.gpu


def test_c5_build_certificate_counts_and_ordered_locate():
    import numpy as np
    import torch
    import csfm_b200 as fm
    from sa_checker import DeviceArrayU32, expected_locate, sa_intervals

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.cuda.empty_cache()
    if torch.cuda.mem_get_info()[0] < 150e9:
        pytest.skip("needs > 150 GB of free device memory")
    n = 4_000_000_000
    text = fm.workloads.dna_text_torch(n, 7, dev)
    torch.cuda.empty_cache()
    idx = fm.FMIndex.build_from_text_device(text.data_ptr(), n, fm.BuildParams(ssa_stride=32), device=0, flags=fm.BUILD_KEEP_SA)
    try:
        info = idx.info()
        assert info.n == n and info.sigma == 5 and info.nsamp == (n + 31) // 32
        cert = fm.workloads.certify_sa_torch(text, idx.sa_device_ptr(), n)
        assert cert["ok"], cert
        torch.cuda.empty_cache()
        sa = DeviceArrayU32.from_ptr(idx.sa_device_ptr(), n, dev)
        # SSA and BWT on sampled rows (the full arrays are 0.5 GB / 4 GB: compared where they are cheap to fetch)
        ssa = idx.ssa()
        rows = torch.arange(0, n, 32 * 4099, dtype=torch.int64, device=dev)
        assert (sa[rows].cpu().numpy() == ssa[::4099].astype(np.int64)).all()
        bwt = torch.from_numpy(idx.bwt()).to(dev)
        rows = fm.workloads._umod_torch_big(fm.workloads.splitmix64_torch(99, torch.arange(1 << 22, dtype=torch.int64, device=dev)), n)
        assert torch.equal(bwt[rows], text[(sa[rows] - 1) % n])
        del bwt
        torch.cuda.empty_cache()

        def run(bytes_d, offs_d, limit_list):
            npat = offs_d.numel() - 1
            lb, ub = sa_intervals(text, sa, bytes_d, offs_d)
            counts = torch.zeros(npat, dtype=torch.int64, device=dev)
            spep = torch.zeros(2 * npat, dtype=torch.int64, device=dev)
            idx.count_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, counts.data_ptr(), 0)
            torch.cuda.synchronize()
            assert torch.equal(counts, ub - lb)
            idx.count_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, counts.data_ptr(), spep.data_ptr())
            torch.cuda.synchronize()
            assert torch.equal(counts, ub - lb)
            hit = ub > lb
            se = spep.view(-1, 2)
            assert torch.equal(se[hit, 0], lb[hit]) and torch.equal(se[hit, 1], ub[hit])
            assert int(se[~hit].abs().sum()) == 0
            for limit in limit_list:
                offs = torch.zeros(npat + 1, dtype=torch.int64, device=dev)
                status = torch.zeros(npat, dtype=torch.int32, device=dev)
                tot = idx.locate_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, limit, offs.data_ptr(), 0, 0, status.data_ptr())
                pos = torch.zeros(max(1, tot), dtype=torch.int64, device=dev)
                idx.locate_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, limit, offs.data_ptr(), pos.data_ptr(), tot,
                                        status.data_ptr())
                torch.cuda.synchronize()
                e_offs, e_pos = expected_locate(sa, lb, ub, limit)
                assert int(status.max()) == 0
                assert torch.equal(offs, e_offs) and torch.equal(pos[:tot], e_pos)
            return counts

        # the config's sweep: text-sampled 20-mers
        b20, o20 = fm.workloads.sampled_patterns_torch(text, 1_000_000, 20, 20, 0, 11)
        c20 = run(b20, o20, [100000, 1])
        assert bool((c20 >= 1).all())
        # random strings: mostly misses, die after ~16 characters
        rd, ro = fm.workloads.random_patterns_np(b"ACGT", 200_000, 20, 5)
        run(torch.from_numpy(rd).to(dev), torch.from_numpy(ro.astype(np.int64)).to(dev), [100000])
        # frequent patterns: length 12 (~240 occurrences each), and the terminator's neighbourhood
        b12, o12 = fm.workloads.sampled_patterns_torch(text, 100_000, 12, 12, 0, 12)
        run(b12, o12, [100000, 5])
        tail = torch.cat([text[n - 21:n], text[n - 1:n], text[0:20]])
        run(tail, torch.tensor([0, 21, 22, 42], dtype=torch.int64, device=dev), [100000])
    finally:
        idx.close()
        del text
        torch.cuda.empty_cache()
