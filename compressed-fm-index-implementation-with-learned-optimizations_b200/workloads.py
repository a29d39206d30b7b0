"""Synthetic workloads of BASELINE.json / SURVEY.md §8d, reproducible across numpy, torch and C.

Everything is a pure function of (seed, index) through splitmix64, never an
implementation-defined distribution (std::uniform_int_distribution is not portable):

    x = seed * 2^32 + i;  z = x + 0x9E3779B97F4A7C15
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9;  z = (z ^ (z >> 27)) * 0x94D049BB133111EB;  z ^= z >> 31

C2  n = 2^26   DNA {A,C,G,T} + '$' terminator, 1 M patterns of length 20 sampled from the text
C3  n = 2^30   bytes 1..255 + 0x00 terminator (sigma = 256, 8 levels), patterns of length 8..32
C4  n = 2^28   DNA as C2, ssa_stride 32, 1 M text-sampled patterns of length 10 / 12 (locate)
"""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1


def splitmix64_np(seed: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (np.uint64((seed << 32) & _M64) + idx.astype(np.uint64)) + np.uint64(0x9E3779B97F4A7C15)
        z = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _lsr(z, k):
    """logical shift right on torch int64"""
    return (z >> k) & ((1 << (64 - k)) - 1)


def splitmix64_torch(seed: int, idx):
    """idx: int64 tensor (values < 2^63). Returns int64 holding the same 64 bits as the numpy version."""
    import torch

    def s64(v):  # python int -> signed 64-bit
        v &= _M64
        return v - (1 << 64) if v >= (1 << 63) else v

    x = idx + s64((seed << 32) + 0x9E3779B97F4A7C15)
    z = (x ^ _lsr(x, 30)) * s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * s64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


# ---- texts ------------------------------------------------------------------------------------
def dna_text_np(n: int, seed: int) -> np.ndarray:
    """ACGT by splitmix64 & 3, last byte '$' (0x24 < 'A': a unique smallest terminator)."""
    t = np.frombuffer(b"ACGT", np.uint8)[(splitmix64_np(seed, np.arange(n, dtype=np.uint64)) & np.uint64(3)).astype(np.int64)]
    t = t.copy()
    t[n - 1] = 0x24
    return t


def byte_text_np(n: int, seed: int) -> np.ndarray:
    """bytes 1..255 uniformly, last byte 0x00 (unique smallest terminator): sigma = 256."""
    t = (np.uint64(1) + splitmix64_np(seed, np.arange(n, dtype=np.uint64)) % np.uint64(255)).astype(np.uint8)
    t[n - 1] = 0
    return t


def _umod_torch(z, m: int):
    """(z as unsigned 64) mod m for torch int64 z, m < 2^31."""
    hi = _lsr(z, 32)
    lo = z & 0xFFFFFFFF
    return ((hi % m) * ((1 << 32) % m) + lo % m) % m


def dna_text_torch(n: int, seed: int, device, chunk: int = 1 << 26):
    import torch
    out = torch.empty(n, dtype=torch.uint8, device=device)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        z = splitmix64_torch(seed, torch.arange(s, e, dtype=torch.int64, device=device))
        out[s:e] = lut[(z & 3)]
    out[n - 1] = 0x24
    return out


def byte_text_torch(n: int, seed: int, device, chunk: int = 1 << 26):
    import torch
    out = torch.empty(n, dtype=torch.uint8, device=device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        z = splitmix64_torch(seed, torch.arange(s, e, dtype=torch.int64, device=device))
        out[s:e] = (1 + _umod_torch(z, 255)).to(torch.uint8)
    out[n - 1] = 0
    return out


# ---- patterns -----------------------------------------------------------------------------------
def sampled_patterns_np(text: np.ndarray, npat: int, len_lo: int, len_hi: int, seed_len: int, seed_pos: int,
                        first: int = 0):
    """Patterns k = first..first+npat-1: len = len_lo + sm(seed_len,k) % (len_hi-len_lo+1), copied from
    text[p : p+len], p = sm(seed_pos,k) % (n - len_hi - 1). Returns (bytes, offs u64[npat+1])."""
    n = text.size
    k = np.arange(first, first + npat, dtype=np.uint64)
    span = len_hi - len_lo + 1
    lens = (len_lo + (splitmix64_np(seed_len, k) % np.uint64(span)).astype(np.int64)) if span > 1 else np.full(npat, len_lo, np.int64)
    pos = (splitmix64_np(seed_pos, k) % np.uint64(n - len_hi - 1)).astype(np.int64)
    offs = np.zeros(npat + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens).astype(np.uint64)
    total = int(offs[-1])
    # gather: byte j of pattern q = text[pos[q] + j]
    q_of = np.repeat(np.arange(npat, dtype=np.int64), lens)
    j_of = np.arange(total, dtype=np.int64) - offs[:-1].astype(np.int64)[q_of]
    data = text[pos[q_of] + j_of]
    return np.ascontiguousarray(data, dtype=np.uint8), offs


def sampled_patterns_torch(text, npat: int, len_lo: int, len_hi: int, seed_len: int, seed_pos: int, first: int = 0):
    """torch twin of sampled_patterns_np on text's device. Returns (bytes u8, offs int64[npat+1])."""
    import torch
    dev = text.device
    n = text.numel()
    k = torch.arange(first, first + npat, dtype=torch.int64, device=dev)
    span = len_hi - len_lo + 1
    if span > 1:
        lens = len_lo + _umod_torch(splitmix64_torch(seed_len, k), span)
    else:
        lens = torch.full((npat,), len_lo, dtype=torch.int64, device=dev)
    pos = _umod_torch_big(splitmix64_torch(seed_pos, k), n - len_hi - 1)
    offs = torch.zeros(npat + 1, dtype=torch.int64, device=dev)
    offs[1:] = torch.cumsum(lens, 0)
    total = int(offs[-1].item())
    q_of = torch.repeat_interleave(torch.arange(npat, dtype=torch.int64, device=dev), lens, output_size=total)
    j_of = torch.arange(total, dtype=torch.int64, device=dev) - offs[:-1][q_of]
    data = text[pos[q_of] + j_of]
    return data.contiguous(), offs


def _umod_torch_big(z, m: int):
    """(z as unsigned 64) mod m for torch int64 z and m up to 2^32: split into 16-bit limbs."""
    r = None
    for shift in (48, 32, 16, 0):
        limb = _lsr(z, shift) & 0xFFFF if shift else z & 0xFFFF
        r = limb % m if r is None else (r * 65536 + limb) % m
    return r


def certify_sa_torch(text, sa_ptr: int, n: int, chunk: int = 1 << 27) -> dict:
    """O(n) certificate, in plain torch on the GPU, that the uint32 array at device pointer `sa_ptr`
    is THE suffix array of `text` in the reference's order (unsigned bytes, a proper prefix first):
    it is a permutation, and for neighbours a = SA[i], b = SA[i+1]: T[a] < T[b], or T[a] == T[b] and
    the suffix after a sorts before the suffix after b (a+1 == n counts as smallest)."""
    import torch
    dev = text.device

    class _Mem:
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    raw = torch.as_tensor(_Mem(sa_ptr, 4 * n), device=dev)

    def sa_chunk(lo, hi):  # uint32 -> int64
        return raw[4 * lo:4 * hi].view(torch.int32).to(torch.int64) & 0xFFFFFFFF

    isa = torch.full((n,), -1, dtype=torch.int64, device=dev)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        s = sa_chunk(lo, hi)
        if int(s.max()) >= n:
            return {"ok": False, "why": "value out of range"}
        isa[s] = torch.arange(lo, hi, dtype=torch.int64, device=dev)
    if int(isa.min()) < 0:
        return {"ok": False, "why": "not a permutation"}
    bad = 0
    for lo in range(0, n - 1, chunk):
        hi = min(n - 1, lo + chunk)
        a, b = sa_chunk(lo, hi), sa_chunk(lo + 1, hi + 1)
        ta, tb = text[a], text[b]
        a1, b1 = a + 1, b + 1
        ra = torch.where(a1 < n, isa[torch.clamp(a1, max=n - 1)], torch.full_like(a, -1))
        rb = torch.where(b1 < n, isa[torch.clamp(b1, max=n - 1)], torch.full_like(b, -1))
        ok = (ta < tb) | ((ta == tb) & (ra < rb))
        bad += int((~ok).sum())
    return {"ok": bad == 0, "violations": bad}


def random_patterns_np(alphabet: bytes, npat: int, length: int, seed: int):
    """Uniform random strings over `alphabet` (mostly misses: die after ~log_sigma(n) steps)."""
    idx = np.arange(npat * length, dtype=np.uint64)
    a = np.frombuffer(alphabet, np.uint8)
    data = a[(splitmix64_np(seed, idx) % np.uint64(len(alphabet))).astype(np.int64)]
    offs = (np.arange(npat + 1, dtype=np.uint64) * np.uint64(length))
    return np.ascontiguousarray(data, dtype=np.uint8), offs
