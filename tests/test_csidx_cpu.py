"""CPU: the `.csidx` container (host/src/serialization/csidx.cpp) against the reference's own
IndexReader (the format oracle; its IndexWriter hangs whenever padding is needed) and,
on the padding-free cases, byte for byte against the reference's IndexWriter."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTLIB = os.path.join(ROOT, "compressed-fm-index-implementation-with-learned-optimizations_b200", "host", "libcs_b200.so")


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(HOSTLIB):
        pytest.skip("libcs_b200.so not built")
    L = C.CDLL(HOSTLIB)
    vp = C.c_void_p
    L.cs_b200_csidx_write.argtypes = [C.c_char_p, C.c_uint32, vp, C.c_uint64, C.c_int, vp, C.c_uint64, vp, C.c_uint64, vp,
                                      C.c_uint64, C.c_uint32, C.c_int, vp, C.c_uint64]
    L.cs_b200_csidx_read.argtypes = [C.c_char_p, vp, vp, vp, vp, vp, vp]
    return L


def _write(L, path, text, bwt, c, ssa, stride, blob, flags=0, has_text=True, has_ssa=True):
    p = lambda a: a.ctypes.data if a.size else None
    rc = L.cs_b200_csidx_write(path.encode(), flags, p(text), text.size, int(has_text), p(bwt), bwt.size, p(c), c.size,
                               p(ssa), ssa.size, stride, int(has_ssa), p(blob), blob.size)
    assert rc == 0


def _read(L, path):
    sizes = np.zeros(8, np.uint64)
    assert L.cs_b200_csidx_read(path.encode(), sizes.ctypes.data, None, None, None, None, None) == 0
    nt, nb, nc, ns, stride, nblob, flags, text_len = [int(x) for x in sizes]
    text, bwt, c, ssa, blob = (np.zeros(max(1, nt), np.uint8), np.zeros(max(1, nb), np.uint8), np.zeros(max(1, nc), np.uint32),
                               np.zeros(max(1, ns), np.uint32), np.zeros(max(1, nblob), np.uint8))
    assert L.cs_b200_csidx_read(path.encode(), sizes.ctypes.data, text.ctypes.data, bwt.ctypes.data, c.ctypes.data,
                                ssa.ctypes.data, blob.ctypes.data) == 0
    return dict(text=text[:nt], bwt=bwt[:nb], c=c[:nc], ssa=ssa[:ns], stride=stride, blob=blob[:nblob], flags=flags,
                text_len=text_len)


def _ref_sections(path):
    R = oracle.ref()
    vp = C.c_void_p
    R.csref_reader_open.restype = vp
    R.csref_reader_open.argtypes = [C.c_char_p]
    R.csref_reader_close.argtypes = [vp]
    R.csref_reader_header.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    R.csref_reader_section.restype = vp
    R.csref_reader_section.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    r = R.csref_reader_open(path.encode())
    assert r, R.csref_last_error()
    flags, text_len, offs = C.c_uint32(), C.c_uint64(), (C.c_uint64 * 8)()
    assert R.csref_reader_header(r, C.byref(flags), C.byref(text_len), offs) == 0
    out = {"flags": flags.value, "text_len": text_len.value, "offsets": list(offs)}
    for which, name, dt in [(1, "text", np.uint8), (2, "bwt", np.uint8), (3, "c", np.uint32), (4, "ssa", np.uint32),
                            (6, "blob", np.uint8)]:
        n, st = C.c_uint64(), C.c_uint32()
        p = R.csref_reader_section(r, which, C.byref(n), C.byref(st))
        out[name] = np.ctypeslib.as_array(C.cast(p, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n.value,)).copy() \
            if p and n.value else np.zeros(0, dt)
        if which == 4:
            out["stride"] = st.value
    R.csref_reader_close(r)
    return out


needs_ref = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/libcsref.so not built")


@needs_ref
@pytest.mark.parametrize("n,stride,blob_len", [(7, 2, 0), (7, 32, 5000), (1000, 7, 4096), (4099, 32, 1), (12345, 16, 70001)])
def test_written_file_reads_back_through_the_reference_reader(host, tmp_path, n, stride, blob_len):
    rng = np.random.default_rng(n)
    text = np.concatenate([rng.integers(1, 200, n - 1, dtype=np.uint8), np.zeros(1, np.uint8)])
    O = oracle.OracleIndex(text, stride=stride)
    blob = rng.integers(0, 256, blob_len, dtype=np.uint8)
    path = str(tmp_path / "t.csidx")
    _write(host, path, text, O.bwt, O.C, O.ssa, stride, blob)
    got = _ref_sections(path)                      # the reference's mmap reader
    assert got["text_len"] == n and (got["text"] == text).all() and (got["bwt"] == O.bwt).all()
    assert (got["c"] == O.C).all() and (got["ssa"] == O.ssa).all() and got["stride"] == stride
    assert (got["blob"] == blob).all()
    assert all(o % 8 == 0 for o in got["offsets"]) and (blob_len == 0 or got["offsets"][6] % 4096 == 0)
    assert bool(got["flags"] & (1 << 8)) == (blob_len > 0)
    mine = _read(host, path)                       # our reader
    for k in ("text", "bwt", "c", "ssa", "blob"):
        assert (mine[k] == got[k]).all()
    assert mine["stride"] == stride and mine["text_len"] == n
    raw = open(path, "rb").read()
    assert raw[:8] == b"CSIDX\0\0\0" and raw[-8:] == bytes.fromhex("0053435345 4e4400".replace(" ", ""))


@needs_ref
def test_byte_identical_to_the_reference_writer_when_no_padding_is_needed(host, tmp_path):
    R = oracle.ref()
    R.csref_writer_simple.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    rng = np.random.default_rng(3)
    text = rng.integers(1, 255, 64, dtype=np.uint8)   # lengths multiple of 8: the reference writer terminates
    bwt = rng.integers(0, 255, 64, dtype=np.uint8)
    c = np.arange(256, dtype=np.uint32) * 3
    a, b = str(tmp_path / "ref.csidx"), str(tmp_path / "ours.csidx")
    assert R.csref_writer_simple(a.encode(), 5, text.ctypes.data, 64, bwt.ctypes.data, 64, c.ctypes.data, 256) == 0
    _write(host, b, text, bwt, c, np.zeros(0, np.uint32), 0, np.zeros(0, np.uint8), flags=5, has_ssa=False)
    ra, rb = open(a, "rb").read(), open(b, "rb").read()
    # ours additionally records an (empty) WAVELET section before the footer: compare everything the
    # reference wrote — header fields it fills, TEXT, BWT, C_ARRAY — and the footer word
    assert rb[:24] == ra[:24]                                  # magic, version, flags, text_len
    ha, hb = np.frombuffer(ra[24:88], np.uint64), np.frombuffer(rb[24:88], np.uint64)
    assert (ha[1:4] == hb[1:4]).all()                          # TEXT / BWT / C_ARRAY offsets
    end_c = int(ha[3]) + 8 + 256 * 4
    assert rb[88:end_c] == ra[88:end_c]                        # the three sections, byte for byte
    assert ra[-8:] == rb[-8:]                                  # footer magic


def test_malformed_files_are_rejected(host, tmp_path):
    text = np.frombuffer(b"banana$", np.uint8)
    path = str(tmp_path / "x.csidx")
    _write(host, path, text, text, np.zeros(257, np.uint32), np.array([6], np.uint32), 32, np.zeros(10, np.uint8))
    raw = bytearray(open(path, "rb").read())
    sizes = np.zeros(8, np.uint64)
    for mutate in (lambda r: r.__setitem__(0, 0x58), lambda r: r.__setitem__(8, 9), lambda r: r.__setitem__(len(r) - 1, 1)):
        bad = bytearray(raw)
        mutate(bad)
        p = str(tmp_path / "bad.csidx")
        open(p, "wb").write(bad)
        assert host.cs_b200_csidx_read(p.encode(), sizes.ctypes.data, None, None, None, None, None) == 1
    open(str(tmp_path / "short.csidx"), "wb").write(raw[:50])
    assert host.cs_b200_csidx_read(str(tmp_path / "short.csidx").encode(), sizes.ctypes.data, None, None, None, None, None) == 1
