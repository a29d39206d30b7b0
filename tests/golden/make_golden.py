#!/usr/bin/env python
"""tests/golden/make_golden.py — regenerates the committed golden fixtures.

Runs ONLY in the build container (needs /root/reference): every expected value below is
produced by the UNMODIFIED reference compiled into oracle/_ref/libcsref.so
(cs::FMIndex::build_from_text / count / locate, cs::WaveletTree, cs::BitVector), never by our
own code. The fixtures travel to the GPU box, where /root/reference does not exist.

  golden_fm.json        small texts: reference test vectors (tests/fm_search_tests.cpp:69-278,
                        tests/simple_tests.cpp, tests/debug_fm.cpp), SURVEY §8a edge cases
                        (no terminator, non-terminating LF walk, 0x00 symbol, empty text),
                        sample.txt / example.txt demo queries (run_all_tests.ps1:45-71),
                        seeded random texts over sigma in {1,2,4,5,256}
  golden_wavelet.json   cs::WaveletTree::rank/access and cs::BitVector::rank1 vectors
                        (tests/wavelet_tests.cpp:63-209, tests/bitvector_tests.cpp:87-187)
  c1_workload.npz       the reference benchmark workload (tools/benchmark.cpp) with per-query
                        reference counts and locate checksums

usage: python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

REF = "/root/reference"


def hx(b) -> str:
    return bytes(b).hex()


def fm_case(name, text: bytes, stride, patterns, limits=(100000,), store_sa=True, source=""):
    R = oracle.RefIndex(text, stride=stride)
    case = {"name": name, "source": source, "text_hex": hx(text), "stride": stride, "n": len(text),
            "bwt_hex": hx(R.bwt.tobytes()), "C": R.C.tolist(), "ssa": R.ssa.tolist()}
    if store_sa:
        case["sa"] = R.sa.tolist()
    qs = []
    for p in patterns:
        p = p.encode("latin-1") if isinstance(p, str) else bytes(p)
        for lim in limits:
            pos, st = R.locate(p, lim)
            qs.append({"pat_hex": hx(p), "limit": lim, "count": R.count(p), "locate": pos, "status": st})
    case["queries"] = qs
    return case


def sampled_patterns(rng, text: bytes, k, lens, sigma_bytes):
    out = []
    n = len(text)
    for _ in range(k):
        m = int(rng.choice(lens))
        if n > m and rng.random() < 0.7:
            s = int(rng.integers(0, n - m))
            out.append(text[s:s + m])
        else:
            out.append(bytes(rng.choice(sigma_bytes, m).astype(np.uint8)))
    return out


def make_fm():
    cases = []
    T = "tests/fm_search_tests.cpp"
    cases.append(fm_case("empty_text", b"", 32, ["", "x", "a"], source=T + ":54-67"))
    cases.append(fm_case("hello", b"hello$", 32, ["", "l", "lo", "hello", "hello$", "$h"], source=T + ":62-64"))
    cases.append(fm_case("banana", b"banana$", 32,
                         ["banana", "ana", "na", "a", "b", "$", "x", "anana", "", "nab", "$b", "a$"],
                         source=T + ":69-113; SURVEY 8a.1"))
    cases.append(fm_case("banana_stride2", b"banana$", 2, ["a", "na", "ana", "banana", "n"], limits=(100000, 2, 1),
                         source="tests/debug_fm.cpp:11-37; SURVEY 8a.1"))
    cases.append(fm_case("no_match", b"abcdefg$", 32, ["xyz", "aaa", "gg", "abc", "g$"], source=T + ":115-128"))
    cases.append(fm_case("multiple", b"aabaabaa$", 4, ["a", "aa", "aab", "b", "baa", "aabaabaa"], source=T + ":130-152"))
    cases.append(fm_case("overlapping", b"abababab$", 32, ["ab", "aba", "abab", "ba", "b$"], source=T + ":154-173"))
    full = bytes(range(1, 256)) + b"$"
    cases.append(fm_case("full_alphabet", full, 32, [bytes([i]) for i in range(0, 256)] + [b"\x01\x02", b"$$"],
                         source=T + ":175-198 (the '$'=0x24 assert at :191 is a known-bad reference test)"))
    pang = (b"The quick brown fox jumps over the lazy dog. The five boxing wizards jump quickly. "
            b"Pack my box with five dozen liquor jugs.$")
    cases.append(fm_case("pangram", pang, 32,
                         ["The", "the", "quick", "fox", "dog", "jump", "five", "box", "xyz", " ", ".", "qu",
                          "ing", "ck", "ox"], source=T + ":237-251"))
    cases.append(fm_case("repeated", b"abcabcabcabc$", 32, ["abc", "ab", "bc", "ca", "abcabc", "a", "c"],
                         source=T + ":253-261"))
    cases.append(fm_case("single_char", b"x$", 32, ["x", "y", "$", "x$", ""], source=T + ":263-278"))
    cases.append(fm_case("mississippi", b"mississippi$", 4, ["ssi", "i", "issi", "p", "s", "mississippi$"],
                         limits=(100000, 3), source="SURVEY 8a.2"))
    # no terminator => cyclic over-count
    cases.append(fm_case("noterm_banana", b"banana", 1, ["anab", "ana", "a", "banana", "nab", "ab"], source="SURVEY 8a.3"))
    cases.append(fm_case("noterm_aaaa_s32", b"aaaa", 32, ["a", "aa", "aaaaa", "b"], source="SURVEY 8a.3/8a.4"))
    cases.append(fm_case("noterm_aaaa_s1", b"aaaa", 1, ["a", "aa", "aaaaa"], source="SURVEY 8a.4"))
    cases.append(fm_case("noterm_abab_s2", b"abab", 2, ["a", "b", "ab", "ba", "abab"], source="SURVEY 8a.4"))
    cases.append(fm_case("noterm_x", b"x", 32, ["", "x", "xx", "y"], source="SURVEY 8a.3/8a.7"))
    cases.append(fm_case("nul_symbol", b"a\x00b\x00", 2, [b"\x00", b"a", b"a\x00", b"\x00b", b"b\x00a"],
                         source="SURVEY 8a.6"))
    # demo files (run_all_tests.ps1:45-71; cs_query does not append a terminator)
    sample = open(os.path.join(REF, "sample.txt"), "rb").read()
    cases.append(fm_case("sample_txt", sample, 32, ["banana", "ana", "band", "an", " ", "$"], limits=(100,),
                         source="sample.txt via tools/query_cli.cpp:11-13"))
    example = open(os.path.join(REF, "example.txt"), "rb").read()
    cases.append(fm_case("example_txt", example, 32,
                         ["algorithm", "quick", "the", "FM-index", "compressed", "The quick brown fox", "\n", "$\n"],
                         limits=(100,), source="example.txt via tools/query_cli.cpp:11-13; SURVEY 8a.5"))
    # seeded random texts
    rng = np.random.default_rng(20261018)
    for sigma, n, stride, term in [(1, 200, 32, True), (2, 300, 4, True), (2, 700, 32, False),
                                   (4, 1000, 32, True), (4, 2049, 1, False), (5, 1500, 8, True),
                                   (256, 2000, 32, True), (256, 4100, 16, False), (17, 3000, 2, True)]:
        if sigma == 256:
            alpha = np.arange(256, dtype=np.uint8)
        else:
            alpha = np.sort(rng.choice(np.arange(1, 256), sigma, replace=False)).astype(np.uint8)
        body = alpha[rng.integers(0, sigma, n)].astype(np.uint8).tobytes()
        text = body + (b"\x00" if term else b"")
        pats = sampled_patterns(rng, text, 40, [1, 2, 3, 4, 6, 9, 14, 25], alpha)
        cases.append(fm_case(f"random_s{sigma}_n{n}_k{stride}_{'t' if term else 'nt'}", text, stride, pats,
                             limits=(100000, 5), store_sa=(n <= 2100), source="seeded numpy default_rng(20261018)"))
    return {"generator": "tests/golden/make_golden.py", "reference": "oracle/_ref/libcsref.so (unmodified cs::FMIndex)",
            "cases": cases}


def make_wavelet():
    rng = np.random.default_rng(42)
    wt_cases = []

    def wt_case(name, seq: bytes, symbols, positions, source):
        W = oracle.RefWavelet(seq)
        ranks = [[int(c), int(i), W.rank(c, i)] for c in symbols for i in positions]
        acc = [W.access(i) for i in range(min(len(seq), 64))]
        return {"name": name, "source": source, "seq_hex": hx(seq), "ranks": ranks, "access_prefix": acc}

    S = "tests/wavelet_tests.cpp"
    wt_cases.append(wt_case("empty", b"", [0, 97], [0, 1, 5], S + ":38-48"))
    wt_cases.append(wt_case("single", b"a", [97, 98], [0, 1, 2], S + ":50-61"))
    wt_cases.append(wt_case("banana", b"banana$", list(b"abn$x"), list(range(0, 9)), S + ":63-91"))
    wt_cases.append(wt_case("zzz", b"z" * 1000, [ord("z"), ord("y")], [0, 1, 500, 999, 1000, 1001], S + ":93-110"))
    for n, seed in [(500, 42), (2000, 123), (5000, 999)]:
        r = np.random.default_rng(seed)
        seq = r.integers(0, 256, n, dtype=np.uint8).tobytes()
        syms = r.integers(0, 256, 8).tolist()
        pos = sorted(set([0, 1, n // 2, n - 1, n] + r.integers(0, n + 1, 16).tolist()))
        wt_cases.append(wt_case(f"random_{n}", seq, syms, pos, S + ":112-151"))
    wt_cases.append(wt_case("all_bytes_twice", bytes(range(256)) * 2, list(range(0, 256, 5)), [0, 1, 255, 256, 257, 511, 512],
                            S + ":153-185"))
    wt_cases.append(wt_case("extremes", bytes([0, 255, 0, 255]), [0, 255, 1], [0, 1, 2, 3, 4], S + ":187-209"))

    bv_cases = []

    def bv_case(name, bits01: np.ndarray, source, positions=None):
        B = oracle.RefBitVector(bits01)
        n = bits01.size
        positions = list(range(0, n + 2)) if positions is None else positions
        packed = np.packbits(bits01, bitorder="little").tobytes()
        return {"name": name, "source": source, "nbits": int(n), "bits_packed_hex": hx(packed),
                "positions": positions, "rank1": [B.rank1(i) for i in positions]}

    Bt = "tests/bitvector_tests.cpp"
    for n in [500, 2048, 5000]:
        bits = rng.integers(0, 2, n, dtype=np.uint8)
        bv_cases.append(bv_case(f"random_{n}", bits, Bt + ":87-124"))
    bits = rng.integers(0, 2, 10000, dtype=np.uint8)
    bv_cases.append(bv_case("random_10000", bits, Bt + ":87-124", positions=list(range(0, 10002, 7)) + [2047, 2048, 2049, 4096, 9999, 10000]))
    for n in [100, 2048, 2049, 5000]:
        bv_cases.append(bv_case(f"zeros_{n}", np.zeros(n, np.uint8), Bt + ":126-160", positions=[0, 1, n // 2, n - 1, n, n + 1]))
        bv_cases.append(bv_case(f"ones_{n}", np.ones(n, np.uint8), Bt + ":126-160", positions=[0, 1, n // 2, n - 1, n, n + 1]))
    bv_cases.append(bv_case("one_bit", np.ones(1, np.uint8), Bt + ":136", positions=[0, 1, 100]))
    # build_from_words 0xAAAA.. / 0x5555.. (bitvector_tests.cpp:171-187)
    for name, w in [("words_aaaa", 0xAAAAAAAAAAAAAAAA), ("words_5555", 0x5555555555555555)]:
        words = np.full(40, w, dtype=np.uint64)
        B = oracle.RefBitVector(words=words, nbits=40 * 64 - 13)
        pos = list(range(0, 40 * 64 - 13 + 2, 3))
        bv_cases.append({"name": name, "source": Bt + ":171-187", "nbits": 40 * 64 - 13, "word": int(w), "nwords": 40,
                         "positions": pos, "rank1": [B.rank1(i) for i in pos]})
    return {"generator": "tests/golden/make_golden.py", "wavelet": wt_cases, "bitvector": bv_cases}


def make_c1():
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "c1gen")
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++20", "-o", exe, os.path.join(HERE, "c1_workload.cpp")], check=True)
        raw = subprocess.run([exe], check=True, capture_output=True).stdout
    n = int(np.frombuffer(raw[:8], np.uint64)[0])
    text = np.frombuffer(raw[8:8 + n], np.uint8).copy()
    npat = int(np.frombuffer(raw[8 + n:16 + n], np.uint64)[0])
    pos = np.frombuffer(raw[16 + n:16 + n + 4 * npat], np.uint32).copy()
    assert n == 100001 and npat == 10000
    print("c1: building the reference index with the verbatim build_from_text (takes ~7 s)...", flush=True)
    R = oracle.RefIndex(text.tobytes(), stride=32)
    rand = [text[p:p + 5].tobytes() for p in pos]
    freq10 = [b"an", b"the", b"ing", b"ed", b"er", b"ba", b"ap", b"or", b"qu", b"la"]
    freq = [freq10[i % 10] for i in range(npat)]
    d, o = oracle.pack_patterns(rand)
    rc = R.count_batch(d, o, nthreads=8)
    d, o = oracle.pack_patterns(freq10)
    fc = R.count_batch(d, o, nthreads=8)
    tot_r, tot_f = int(rc.sum()), int(fc.sum()) * (npat // 10)
    print("c1: total_matches random/frequent =", tot_r, tot_f)
    # SURVEY §6 probe of the unmodified tools/benchmark.cpp in this toolchain
    assert tot_r == 9907582, tot_r
    assert tot_f == 16309000, tot_f
    # locate benchmark: first 100 frequent patterns, default limit (benchmark.cpp:173-224, :330-336)
    loc_n, loc_sum, loc_xor, loc_first = [], [], [], []
    for q in freq10:
        p, st = R.locate(q, 100000)
        assert st == 0
        a = np.array(p, dtype=np.uint64)
        loc_n.append(a.size)
        loc_sum.append(int(a.sum()))
        loc_xor.append(int(np.bitwise_xor.reduce(a)) if a.size else 0)
        loc_first.append((a[:8].tolist() + [0] * 8)[:8])
    assert sum(loc_n) * 10 == 163090, sum(loc_n) * 10
    np.savez_compressed(os.path.join(HERE, "c1_workload.npz"), text=text, rand_pos=pos, rand_count=rc.astype(np.uint32),
                        freq_patterns=np.array([p.ljust(3, b"\0") for p in freq10], dtype="S3"),
                        freq_len=np.array([len(p) for p in freq10], np.uint8), freq_count=fc.astype(np.uint32),
                        sa=R.sa, ssa=R.ssa, C=R.C,
                        loc_n=np.array(loc_n, np.uint64), loc_sum=np.array(loc_sum, np.uint64),
                        loc_xor=np.array(loc_xor, np.uint64), loc_first=np.array(loc_first, np.uint64))


if __name__ == "__main__":
    oracle.build()
    assert oracle.ref_available(), "needs /root/reference"
    json.dump(make_fm(), open(os.path.join(HERE, "golden_fm.json"), "w"), separators=(",", ":"))
    json.dump(make_wavelet(), open(os.path.join(HERE, "golden_wavelet.json"), "w"), separators=(",", ":"))
    make_c1()
    for f in ["golden_fm.json", "golden_wavelet.json", "c1_workload.npz"]:
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
