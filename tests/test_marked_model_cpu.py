"""CPU model of the marked line form of layout 3 (csrc/csfm_dna.cuh) — test infrastructure, not the product.

It restates in numpy what `dna_pack_kernel<true>` writes and what `walk3_kernel<., true>` does with it, and checks the
argument the form rests on against the oracle (the reference's own SA / BWT): when the LAST byte of the text occurs
exactly once — whatever its value, wherever it sorts — LF maps the row of suffix j to the row of suffix j - 1 for every
j > 0, so a walk that stops at the first row whose suffix starts at a multiple of the stride returns SA[row] after
exactly SA[row] mod stride steps. The GPU tests compare the real lines with this model word by word
(tests/test_gpu_parity.py::test_marked_lines_build_products)."""
import numpy as np
import pytest

import oracle


def marked_lines(codes, sa, n, stride):
    """codes: two-bit code per BWT row (the symbol without a code stored as 0). -> (lines u32[nblk, 16], psamp u32[])"""
    nblk = n // 128 + 1
    sym = np.zeros(nblk * 128, np.int64); sym[:n] = codes
    mark = np.zeros(nblk * 128, np.int64); mark[:n] = sa % stride == 0
    valid = np.zeros(nblk * 128, bool); valid[:n] = True
    w = 1 << np.arange(32, dtype=np.uint64)
    words = lambda bits: (bits.reshape(nblk, 4, 32).astype(np.uint64) * w).sum(axis=2).astype(np.uint32)
    before = lambda bits: (np.cumsum(bits.reshape(nblk, 128).sum(axis=1)) - bits.reshape(nblk, 128).sum(axis=1)).astype(np.uint32)
    lines = np.zeros((nblk, 16), np.uint32)
    lo, hi, mk = words(sym & 1), words(sym >> 1), words(mark)
    for t in range(4):
        base = 8 * (t >> 1)
        lines[:, base + 2 + 2 * (t & 1)], lines[:, base + 3 + 2 * (t & 1)], lines[:, base + 6 + (t & 1)] = lo[:, t], hi[:, t], mk[:, t]
    lines[:, 0], lines[:, 1], lines[:, 8], lines[:, 9] = before((sym == 0) & valid), before(sym == 1), before(sym == 2), before(mark)
    return lines, sa[sa % stride == 0].astype(np.uint32)


def line_query(lines, p):
    """What one fetch of line p // 128 yields: (symbol at p, rank of that symbol before p, mark bit of p, marks before p)."""
    b, off = p // 128, p % 128
    ln = lines[b]
    word = lambda kind, t: int(ln[8 * (t >> 1) + {"lo": 2, "hi": 3}[kind] + 2 * (t & 1)]) if kind != "mk" else int(ln[8 * (t >> 1) + 6 + (t & 1)])
    t, s = off // 32, off % 32
    v = ((word("lo", t) >> s) & 1) | (((word("hi", t) >> s) & 1) << 1)
    below = lambda x, tt: bin(x & ((1 << max(0, min(32, off - 32 * tt))) - 1)).count("1")
    hits = 0
    for tt in range(4):
        lo_w, hi_w = word("lo", tt), word("hi", tt)
        m = (lo_w if v & 1 else ~lo_w) & (hi_w if v & 2 else ~hi_w) & 0xFFFFFFFF
        hits += below(m, tt)
    c0, c1, c2 = int(ln[0]), int(ln[1]), int(ln[8])
    cnt = [c0, c1, c2, 128 * b - c0 - c1 - c2][v]
    return v, cnt + hits, (word("mk", t) >> s) & 1, int(ln[9]) + sum(below(word("mk", tt), tt) for tt in range(4))


@pytest.mark.parametrize("letters,n,stride,last", [(4, 700, 8, 0x24), (4, 513, 5, 0xFF), (4, 384, 32, 0x50), (3, 300, 3, 0x42),
                                                   (2, 260, 7, 0x00), (1, 130, 4, 0x7A), (4, 129, 1, 0x24), (4, 200, 1000, 0x24)])
def test_marked_walk_model_returns_sa(letters, n, stride, last):
    rng = np.random.default_rng(n * 31 + stride)
    alpha = np.array([0x41, 0x43, 0x47, 0x54][:letters], np.uint8)
    text = np.concatenate([alpha[rng.integers(0, letters, n - 1)], np.array([last], np.uint8)]).astype(np.uint8)
    orc = oracle.OracleIndex(text, stride=stride)
    sa, bwt, C = orc.sa.astype(np.int64), orc.bwt, orc.C.astype(np.int64)
    present = np.unique(text)
    special = last if present.size == 5 else None       # five symbols: the one that occurs once has no two-bit code
    coded = present[present != last] if special is not None else present
    code = np.zeros(256, np.int64); code[coded] = np.arange(coded.size)
    lines, psamp = marked_lines(code[bwt], sa, n, stride)
    assert psamp.size == (n + stride - 1) // stride
    px = int(np.flatnonzero(bwt == last)[0]) if special is not None else -1
    base = {int(code[b]): int(C[b]) for b in coded}
    for row in range(n):
        p, steps = row, 0
        while True:
            v, r, mk, mr = line_query(lines, p)
            if mk:
                break
            # LF(p) = C[c] + occ(c, p); the symbol without a code sits at row px, is stored as a 0 and counted as one
            p = int(C[last]) if p == px else base[v] + r - (1 if (v == 0 and px >= 0 and p > px) else 0)
            steps += 1
            assert steps < stride and steps < n
        assert int(psamp[mr]) + steps == sa[row]
        assert steps == sa[row] % stride
