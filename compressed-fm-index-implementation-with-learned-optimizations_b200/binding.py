"""ctypes binding of libcsfm.so (include/csfm.h) plus a Python mirror of cs::FMIndex.

The product's host side is C++ (host/src/api/fm_index.hpp); this module exists so that the
parity tests and bench.py can drive the SAME C ABI from Python. It mirrors the reference
interface (/root/reference/src/api/fm_index.hpp:11-37): ``BuildParams``, ``FMIndex.build_from_text``,
``count``, ``locate`` (raises RuntimeError with the reference's message where the reference
throws), ``extract``, and adds ``count_batch`` / ``locate_batch``.

There is no CPU fallback: importing works anywhere (so the symbol-export test can run without
a GPU), but every compute call fails loudly when the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CSFM_LIB") or os.path.join(HERE, "libcsfm.so")  # CSFM_LIB: an experiment build of the same sources

CSFM_OK, CSFM_ERR_INVALID, CSFM_ERR_CUDA, CSFM_ERR_NOMEM, CSFM_ERR_TOO_LARGE, CSFM_ERR_CAPACITY, CSFM_ERR_FORMAT = range(7)
Q_OK, Q_LF_WALK_EXCEEDED, Q_SSA_OOB = 0, 1, 2
BUILD_DEFAULT, BUILD_NO_COMPACT, BUILD_KEEP_SA, BUILD_LAYOUT_BINARY64, BUILD_NO_KMER_TABLE, BUILD_NO_TEXT_CHECK, BUILD_FORCE_TEXT_CHECK = 0, 1, 2, 4, 8, 16, 32
BUILD_LARGE_TABLE = 64
BUILD_LAYOUT_NIBBLE128 = 128
BUILD_ROW_SAMPLES = 256

LF_WALK_MESSAGE = "locate: LF walk exceeded text length"  # fm_index.cpp:137

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p


class CsfmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"csfm error {code}: {msg}")
        self.code = code


class Params(C.Structure):  # csfm_params == cs::BuildParams
    _fields_ = [("S", C.c_uint32), ("s", C.c_uint32), ("ssa_stride", C.c_uint32), ("eps", C.c_double)]


class IndexInfo(C.Structure):
    _fields_ = [("n", C.c_uint64), ("sigma", C.c_uint32), ("levels", C.c_uint32), ("ssa_stride", C.c_uint32),
                ("device", C.c_uint32), ("nsamp", C.c_uint64), ("blocks_per_level", C.c_uint64),
                ("blob_bytes", C.c_uint64), ("has_sa", C.c_uint32), ("layout", C.c_uint32), ("line_bytes", C.c_uint32),
                ("kmer_k", C.c_uint32), ("text_check", C.c_uint32), ("half_table", C.c_uint32),
                ("sa_rounds", C.c_uint32), ("sa_radix_passes", C.c_uint32), ("sa_pair_passes", C.c_uint64),
                ("position_samples", C.c_uint32), ("pad0", C.c_uint32)]


class CallStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("search_steps", C.c_uint64), ("lf_steps", C.c_uint64), ("kernel_ms", C.c_float),
                ("table_lookups", C.c_uint32), ("text_checks", C.c_uint32), ("half_steps", C.c_uint32), ("pad0", C.c_uint32),
                ("line_fetches", C.c_uint64)]


# name -> (restype, argtypes): every symbol include/csfm.h declares
SIGNATURES = {
    "csfm_last_error": (C.c_char_p, []),
    "csfm_version": (C.c_char_p, []),
    "csfm_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "csfm_build_from_text": (C.c_int, [_vp, C.c_uint64, C.POINTER(Params), C.c_int, C.c_uint32, C.POINTER(_vp)]),
    "csfm_build_from_text_device": (C.c_int, [_vp, C.c_uint64, C.POINTER(Params), C.c_int, C.c_uint32, C.POINTER(_vp)]),
    "csfm_build_from_parts": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32, C.POINTER(_vp)]),
    "csfm_destroy": (None, [_vp]),
    "csfm_info": (C.c_int, [_vp, C.POINTER(IndexInfo)]),
    "csfm_get_C": (C.c_int, [_vp, _vp]),
    "csfm_get_ssa": (C.c_int, [_vp, _vp]),
    "csfm_get_sa": (C.c_int, [_vp, _vp]),
    "csfm_sa_device": (C.c_int, [_vp, C.POINTER(_vp)]),
    "csfm_release_sa": (C.c_int, [_vp]),
    "csfm_extract_bwt": (C.c_int, [_vp, _vp]),
    "csfm_extract": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp, C.POINTER(C.c_uint64)]),
    "csfm_blob": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_uint64)]),
    "csfm_attach_blob": (C.c_int, [_vp, C.c_uint64, C.c_int, C.c_int, C.POINTER(_vp)]),
    "csfm_alias": (C.c_int, [_vp, C.POINTER(_vp)]),
    "csfm_replicate": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "csfm_blob_to_host": (C.c_int, [_vp, _vp, C.c_uint64]),
    "csfm_from_host_blob": (C.c_int, [_vp, C.c_uint64, C.c_int, C.POINTER(_vp)]),
    "csfm_count_batch": (C.c_int, [_vp, _vp, _vp, C.c_uint64, _vp, _vp]),
    "csfm_count_batch_device": (C.c_int, [_vp, _vp, _vp, C.c_uint64, _vp, _vp, _vp]),
    "csfm_count_batch_submit": (C.c_int, [_vp, _vp, _vp, C.c_uint64, _vp, _vp, C.POINTER(C.c_uint64)]),
    "csfm_count_batch_submit32": (C.c_int, [_vp, _vp, _vp, C.c_uint64, _vp, C.POINTER(C.c_uint64)]),
    "csfm_count_batch_submit_len8": (C.c_int, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _vp, C.POINTER(C.c_uint64)]),
    "csfm_pattern_codes": (C.c_int, [_vp, _vp, C.POINTER(C.c_uint32)]),
    "csfm_count_batch_submit_packed": (C.c_int, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _vp, C.POINTER(C.c_uint64)]),
    "csfm_count_batch_wait": (C.c_int, [_vp, C.c_uint64]),
    "csfm_locate_batch": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp, _vp, C.c_uint64, _vp, C.POINTER(C.c_uint64)]),
    "csfm_locate_batch_device": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp, _vp, C.c_uint64, _vp,
                                           C.POINTER(C.c_uint64), _vp]),
    "csfm_set_instrumentation": (C.c_int, [_vp, C.c_uint32]),
    "csfm_last_call_stats": (C.c_int, [_vp, C.POINTER(CallStats)]),
    "csfm_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_uint64]),
    "csfm_host_free": (C.c_int, [_vp]),
}

_lib = None


def lib():
    """Loads libcsfm.so. Raises (never falls back) when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                              "(there is no CPU fallback for the FM-index engine)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != CSFM_OK:
        raise CsfmError(rc, lib().csfm_last_error().decode(errors="replace"))


def _as_u8(data) -> np.ndarray:
    if isinstance(data, str):
        data = data.encode("latin-1")
    if isinstance(data, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(data), dtype=np.uint8)
    return np.ascontiguousarray(data, dtype=np.uint8)


def _np_ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def pack_patterns(patterns):
    """list of bytes/str -> (bytes u8[total], offs u64[npat+1])"""
    arrs = [_as_u8(p) for p in patterns]
    offs = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        offs[1:] = np.cumsum([a.size for a in arrs], dtype=np.uint64)
    data = np.concatenate(arrs) if arrs and int(offs[-1]) else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(data, dtype=np.uint8), offs


@dataclass
class BuildParams:
    """cs::BuildParams (fm_index.hpp:11-14). Only ssa_stride is honoured, as in the reference."""
    S: int = 512
    s: int = 64
    ssa_stride: int = 32
    eps: float = 1.0

    def c(self) -> Params:
        return Params(self.S, self.s, self.ssa_stride, self.eps)


class FMIndex:
    """Device-resident FM-index with the reference's cs::FMIndex interface."""

    def __init__(self, handle, text=None):
        self._h = _vp(handle) if not isinstance(handle, _vp) else handle
        self._text = text  # kept on the host for extract(), like FMIndex::text_ (fm_index.hpp:41)
        self._keepalive = None

    # ---- construction --------------------------------------------------------------------
    @staticmethod
    def build_from_text(text, params: BuildParams | None = None, device: int = 0, flags: int = 0) -> "FMIndex":
        """cs::FMIndex::build_from_text (fm_index.cpp:16-69): SA, BWT, C, wavelet, SSA on the GPU."""
        t = _as_u8(text)
        p = (params or BuildParams()).c()
        h = _vp()
        _check(lib().csfm_build_from_text(_np_ptr(t), t.size, C.byref(p), device, flags, C.byref(h)))
        return FMIndex(h, text=t.tobytes())

    @staticmethod
    def build_from_text_device(d_text_ptr: int, n: int, params: BuildParams | None = None, device: int = 0,
                               flags: int = 0) -> "FMIndex":
        p = (params or BuildParams()).c()
        h = _vp()
        _check(lib().csfm_build_from_text_device(_vp(d_text_ptr), n, C.byref(p), device, flags, C.byref(h)))
        return FMIndex(h)

    @staticmethod
    def from_parts(bwt, ssa, ssa_stride: int, device: int = 0, flags: int = 0, text=None) -> "FMIndex":
        b = _as_u8(bwt)
        s = np.ascontiguousarray(ssa, dtype=np.uint32)
        h = _vp()
        _check(lib().csfm_build_from_parts(_np_ptr(b), b.size, _np_ptr(s), s.size, ssa_stride, device, flags, C.byref(h)))
        return FMIndex(h, text=None if text is None else _as_u8(text).tobytes())

    @staticmethod
    def open_directory(_dir: str) -> "FMIndex":
        raise RuntimeError("on-disk open not implemented yet")  # fm_index.cpp:71-73

    @staticmethod
    def attach_blob(d_ptr: int, nbytes: int, device: int = 0, keepalive=None) -> "FMIndex":
        h = _vp()
        _check(lib().csfm_attach_blob(_vp(d_ptr), nbytes, device, 0, C.byref(h)))
        idx = FMIndex(h)
        idx._keepalive = keepalive
        return idx

    def alias(self) -> "FMIndex":
        """A second handle over the same device blob (no copy) for another host thread (csfm_alias). Keeps this
        handle alive for as long as the alias lives."""
        h = _vp()
        _check(lib().csfm_alias(self._h, C.byref(h)))
        a = FMIndex(h, text=self._text)
        a._keepalive = self
        return a

    def replicate(self, device: int) -> "FMIndex":
        """A second handle over its own copy of the blob on `device` (csfm_replicate)."""
        h = _vp()
        _check(lib().csfm_replicate(self._h, int(device), C.byref(h)))
        return FMIndex(h, text=self._text)

    @staticmethod
    def from_host_blob(blob: np.ndarray, device: int = 0) -> "FMIndex":
        b = np.ascontiguousarray(blob, dtype=np.uint8)
        h = _vp()
        _check(lib().csfm_from_host_blob(_np_ptr(b), b.size, device, C.byref(h)))
        return FMIndex(h)

    def close(self):
        if self._h:
            lib().csfm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection ---------------------------------------------------------------------
    def info(self) -> IndexInfo:
        i = IndexInfo()
        _check(lib().csfm_info(self._h, C.byref(i)))
        return i

    @property
    def n(self) -> int:
        return int(self.info().n)

    def C_array(self) -> np.ndarray:
        out = np.zeros(257, np.uint32)
        _check(lib().csfm_get_C(self._h, _np_ptr(out)))
        return out

    def ssa(self) -> np.ndarray:
        out = np.zeros(max(1, int(self.info().nsamp)), np.uint32)
        _check(lib().csfm_get_ssa(self._h, _np_ptr(out)))
        return out[: int(self.info().nsamp)]

    def sa(self) -> np.ndarray:
        out = np.zeros(max(1, self.n), np.uint32)
        _check(lib().csfm_get_sa(self._h, _np_ptr(out)))
        return out[: self.n]

    def sa_device_ptr(self) -> int:
        p = _vp()
        _check(lib().csfm_sa_device(self._h, C.byref(p)))
        return int(p.value or 0)

    def release_sa(self):
        _check(lib().csfm_release_sa(self._h))

    def bwt(self) -> np.ndarray:
        out = np.zeros(max(1, self.n), np.uint8)
        _check(lib().csfm_extract_bwt(self._h, _np_ptr(out)))
        return out[: self.n]

    def blob(self):
        """-> (device pointer, bytes) of the contiguous index blob."""
        p, b = _vp(), C.c_uint64()
        _check(lib().csfm_blob(self._h, C.byref(p), C.byref(b)))
        return int(p.value), int(b.value)

    def blob_to_host(self) -> np.ndarray:
        _, nb = self.blob()
        out = np.zeros(nb, np.uint8)
        _check(lib().csfm_blob_to_host(self._h, _np_ptr(out), nb))
        return out

    def set_instrumentation(self, mask: int):
        _check(lib().csfm_set_instrumentation(self._h, mask))

    def last_call_stats(self) -> CallStats:
        s = CallStats()
        _check(lib().csfm_last_call_stats(self._h, C.byref(s)))
        return s

    # ---- batched queries (host buffers) ------------------------------------------------------
    def count_batch(self, data, offs, want_intervals=False):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        npat = offs.size - 1
        counts = np.zeros(max(1, npat), np.uint64)
        sp_ep = np.zeros(max(1, 2 * npat), np.uint64) if want_intervals else None
        _check(lib().csfm_count_batch(self._h, _np_ptr(data), _np_ptr(offs), npat, _np_ptr(counts),
                                      _np_ptr(sp_ep) if want_intervals else None))
        if want_intervals:
            return counts[:npat], sp_ep[: 2 * npat].reshape(-1, 2)
        return counts[:npat]

    def locate_batch(self, data, offs, limit=100000):
        """-> (out_offs u64[npat+1], positions u64[total], status i32[npat])"""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        npat = offs.size - 1
        out_offs = np.zeros(npat + 1, np.uint64)
        status = np.zeros(max(1, npat), np.int32)
        total = C.c_uint64()
        L = lib()
        _check(L.csfm_locate_batch(self._h, _np_ptr(data), _np_ptr(offs), npat, int(limit), _np_ptr(out_offs), None, 0,
                                   _np_ptr(status), C.byref(total)))
        pos = np.zeros(max(1, total.value), np.uint64)
        if total.value:
            _check(L.csfm_locate_batch(self._h, _np_ptr(data), _np_ptr(offs), npat, int(limit), _np_ptr(out_offs),
                                       _np_ptr(pos), total.value, _np_ptr(status), C.byref(total)))
        return out_offs, pos[: total.value], status[:npat]

    # ---- device-pointer variants (torch tensors or raw ints) --------------------------------------
    def count_batch_device(self, d_bytes: int, d_offs: int, npat: int, d_counts: int, d_sp_ep: int = 0, stream: int = 0):
        _check(lib().csfm_count_batch_device(self._h, _vp(d_bytes), _vp(d_offs), npat, _vp(d_counts),
                                             _vp(d_sp_ep) if d_sp_ep else None, _vp(stream) if stream else None))

    def locate_batch_device(self, d_bytes: int, d_offs: int, npat: int, limit: int, d_out_offs: int, d_out_pos: int,
                            cap: int, d_status: int, stream: int = 0) -> int:
        total = C.c_uint64()
        _check(lib().csfm_locate_batch_device(self._h, _vp(d_bytes), _vp(d_offs), npat, int(limit), _vp(d_out_offs),
                                              _vp(d_out_pos) if d_out_pos else None, cap,
                                              _vp(d_status) if d_status else None, C.byref(total),
                                              _vp(stream) if stream else None))
        return int(total.value)

    def count_batch_submit(self, bytes_ptr: int, offs_ptr: int, npat: int, counts_ptr: int, sp_ep_ptr: int = 0) -> int:
        """Asynchronous count over raw (pinned) host pointers; returns a ticket for count_batch_wait."""
        t = C.c_uint64()
        _check(lib().csfm_count_batch_submit(self._h, _vp(bytes_ptr), _vp(offs_ptr), npat, _vp(counts_ptr),
                                             _vp(sp_ep_ptr) if sp_ep_ptr else None, C.byref(t)))
        return int(t.value)

    def count_batch_submit32(self, bytes_ptr: int, offs32_ptr: int, npat: int, counts32_ptr: int) -> int:
        """Compact asynchronous count: u32 offsets in, u32 counts out (less PCIe traffic per query)."""
        t = C.c_uint64()
        _check(lib().csfm_count_batch_submit32(self._h, _vp(bytes_ptr), _vp(offs32_ptr), npat, _vp(counts32_ptr), C.byref(t)))
        return int(t.value)

    def count_batch_submit_len8(self, bytes_ptr: int, nbytes: int, lens8_ptr: int, npat: int, counts32_ptr: int) -> int:
        """Most compact asynchronous count: one length byte per pattern in, u32 counts out."""
        t = C.c_uint64()
        _check(lib().csfm_count_batch_submit_len8(self._h, _vp(bytes_ptr), nbytes, _vp(lens8_ptr), npat, _vp(counts32_ptr),
                                                  C.byref(t)))
        return int(t.value)

    def pattern_codes(self):
        """-> (code_of_byte u8[256], bits): the wire codes of csfm_count_batch_submit_packed."""
        codes = np.zeros(256, np.uint8)
        bits = C.c_uint32()
        _check(lib().csfm_pattern_codes(self._h, _np_ptr(codes), C.byref(bits)))
        return codes, int(bits.value)

    def pack_codes(self, data) -> np.ndarray:
        """Pattern bytes -> the packed wire form (LSB-first, `bits` per symbol, back to back)."""
        codes, bits = self.pattern_codes()
        sym = codes[np.ascontiguousarray(data, dtype=np.uint8)]
        planes = ((sym[:, None] >> np.arange(bits, dtype=np.uint8)[None, :]) & 1).astype(np.uint8)
        return np.packbits(planes.reshape(-1), bitorder="little")

    def count_batch_submit_packed(self, packed_ptr: int, nsyms: int, lens8_ptr: int, npat: int, counts32_ptr: int) -> int:
        """Smallest asynchronous count: packed wire codes + one length byte per pattern in, u32 counts out."""
        t = C.c_uint64()
        _check(lib().csfm_count_batch_submit_packed(self._h, _vp(packed_ptr), nsyms, _vp(lens8_ptr), npat, _vp(counts32_ptr),
                                                    C.byref(t)))
        return int(t.value)

    def count_batch_wait(self, ticket: int):
        _check(lib().csfm_count_batch_wait(self._h, ticket))

    # ---- the reference's single-query API -----------------------------------------------------
    def count(self, pattern) -> int:
        """cs::FMIndex::count (fm_index.cpp:79-101)."""
        d, o = pack_patterns([pattern])
        return int(self.count_batch(d, o)[0])

    def locate(self, pattern, limit: int = 100000):
        """cs::FMIndex::locate (fm_index.cpp:107-157): SA-row order, raises where it throws."""
        d, o = pack_patterns([pattern])
        offs, pos, status = self.locate_batch(d, o, limit)
        if status[0] == Q_LF_WALK_EXCEEDED:
            raise RuntimeError(LF_WALK_MESSAGE)
        if status[0] == Q_SSA_OOB:
            raise RuntimeError("locate: SSA sample index out of range")
        return pos.tolist()

    def extract(self, pos: int, length: int) -> bytes:
        """cs::FMIndex::extract (fm_index.cpp:163-167): host substring, clamped."""
        if self._text is None:  # no host copy (attached / loaded blob): from the index itself
            out = np.zeros(max(1, int(length)), np.uint8)
            got = C.c_uint64()
            _check(lib().csfm_extract(self._h, int(pos), int(length), _np_ptr(out), C.byref(got)))
            return out[: got.value].tobytes()
        if pos >= len(self._text):
            return b""
        return self._text[pos: pos + min(length, len(self._text) - pos)]


def host_alloc(nbytes: int) -> int:
    p = _vp()
    _check(lib().csfm_host_alloc(C.byref(p), nbytes))
    return int(p.value)


def host_free(ptr: int):
    lib().csfm_host_free(_vp(ptr))
