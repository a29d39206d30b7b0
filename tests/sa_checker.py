"""Independent full-batch checker for count / locate at sizes no CPU index can reach (TEST infrastructure,
never imported by the product).

For a text that ends in a unique smallest byte the rows of the reference's BWT matrix are the sorted
suffixes (sais.hpp:8-16), so the interval the reference's backward search ends with
(fm_index.cpp:79-101) is exactly [lower_bound, upper_bound) of the pattern among the suffixes, and
locate (fm_index.cpp:107-157) reports SA[sp], SA[sp+1], ... in that order. Given a CERTIFIED suffix
array (workloads.certify_sa_torch: a certified SA is THE reference SA) this module finds both bounds
by binary search over SA with direct comparisons against the text, in plain torch: nothing of the
engine's index (levels, tables, kernels) takes part.

Works on any torch device (the CPU tests pin it against the oracle and the compiled reference).
"""
from __future__ import annotations


def sa_as_int64(sa_u32_bytes_view, lo=None, hi=None):
    """uint8 view of a uint32 array -> int64 values (torch has no uint32 arithmetic)."""
    import torch
    v = sa_u32_bytes_view.view(torch.int32)
    if lo is not None:
        v = v[lo:hi]
    return v.to(torch.int64) & 0xFFFFFFFF


class DeviceArrayU32:
    """A uint32 device array (raw pointer or uint8/int32 tensor) gathered as int64."""

    def __init__(self, tensor_i32):
        self.t = tensor_i32

    @staticmethod
    def from_ptr(ptr: int, count: int, device):
        import torch

        class _Mem:
            def __init__(self, p, nbytes):
                self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (p, False), "version": 2}

        return DeviceArrayU32(torch.as_tensor(_Mem(ptr, 4 * count), device=device).view(torch.int32))

    def __getitem__(self, idx):
        import torch
        return self.t[idx].to(torch.int64) & 0xFFFFFFFF


def sa_intervals(text, sa: DeviceArrayU32, bytes_d, offs_d, chunk: int = 1 << 17):
    """-> (lb, ub) int64[npat]: rows [lb, ub) of the suffix array whose suffixes start with the pattern.
    Empty patterns get (0, n). Comparison order: unsigned bytes, a suffix that ends inside the pattern
    sorts before it (std::string::operator<, sais.hpp:13)."""
    import torch
    dev = text.device
    n = text.numel()
    npat = offs_d.numel() - 1
    lens_all = offs_d[1:] - offs_d[:-1]
    lb = torch.zeros(npat, dtype=torch.int64, device=dev)
    ub = torch.zeros(npat, dtype=torch.int64, device=dev)
    if npat == 0:
        return lb, ub
    if n == 0:
        return lb, ub
    for c0 in range(0, npat, chunk):
        c1 = min(npat, c0 + chunk)
        lens = lens_all[c0:c1]
        m = int(lens.max())
        k = torch.arange(max(m, 1), dtype=torch.int64, device=dev)[None, :]
        live = k < lens[:, None]
        src = torch.where(live, offs_d[c0:c1, None] + k, torch.zeros_like(k))
        P = torch.where(live, bytes_d[src].to(torch.int16), torch.full_like(src, -2, dtype=torch.int16))
        for upper in (False, True):
            lo = torch.zeros(c1 - c0, dtype=torch.int64, device=dev)
            hi = torch.full((c1 - c0,), n, dtype=torch.int64, device=dev)
            for _ in range(34):
                open_ = lo < hi
                if not bool(open_.any()):
                    break
                mid = (lo + hi) >> 1
                s = sa[torch.clamp(mid, max=n - 1)]
                ti = s[:, None] + k
                inside = ti < n
                T = torch.where(inside, text[torch.clamp(ti, max=n - 1)].to(torch.int16),
                                torch.full_like(ti, -1, dtype=torch.int16))  # past the end: below every byte
                diff = (T != P) & live
                has = diff.any(dim=1)
                first = torch.argmax(diff.to(torch.uint8), dim=1)
                t_first = torch.gather(T, 1, first[:, None]).squeeze(1)
                p_first = torch.gather(P, 1, first[:, None]).squeeze(1)
                less = has & (t_first < p_first)         # suffix[:m] <  pattern
                go_right = (less | ~has) if upper else less  # upper bound: suffix[:m] <= pattern
                lo = torch.where(open_ & go_right, mid + 1, lo)
                hi = torch.where(open_ & ~go_right, mid, hi)
            (ub if upper else lb)[c0:c1] = lo
    return lb, ub


def expected_locate(sa: DeviceArrayU32, lb, ub, limit: int, lens=None):
    """-> (out_offs int64[npat+1], positions int64[total]) the reference's locate would report:
    SA[lb + k] for k < min(ub - lb, limit) (fm_index.cpp:118-152); nothing for empty patterns."""
    import torch
    cnt = torch.clamp(ub - lb, max=limit)
    if lens is not None:
        cnt = torch.where(lens > 0, cnt, torch.zeros_like(cnt))
    offs = torch.zeros(cnt.numel() + 1, dtype=torch.int64, device=cnt.device)
    offs[1:] = torch.cumsum(cnt, 0)
    total = int(offs[-1])
    if total == 0:
        return offs, torch.zeros(0, dtype=torch.int64, device=cnt.device)
    q_of = torch.repeat_interleave(torch.arange(cnt.numel(), device=cnt.device), cnt, output_size=total)
    k = torch.arange(total, device=cnt.device) - offs[:-1][q_of]
    pos = torch.empty(total, dtype=torch.int64, device=cnt.device)
    step = 1 << 26
    for a in range(0, total, step):
        b = min(total, a + step)
        pos[a:b] = sa[lb[q_of[a:b]] + k[a:b]]
    return offs, pos
