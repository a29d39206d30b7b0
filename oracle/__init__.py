"""oracle — TEST INFRASTRUCTURE ONLY.

ctypes bindings for the two CPU checkers:

* ``OracleIndex``  -> ``oracle/liboracle.so``  (fm_oracle.c, the plain-C restatement)
* ``RefIndex``     -> ``oracle/_ref/libcsref.so`` (the UNMODIFIED reference TUs compiled from
  /root/reference by ``oracle/Makefile`` + the forwarding shim ref_harness.cpp)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module. The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcsref.so")

ORC_OK, ORC_LF_WALK_EXCEEDED, ORC_SSA_OOB = 0, 1, 2

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


def build(force: bool = False) -> None:
    """Compile the checkers (``make -C oracle``). Building the checker is not using it."""
    if force or not os.path.exists(ORACLE_SO) or (
        os.path.exists("/root/reference/src/api/fm_index.cpp") and not os.path.exists(REF_SO)
    ):
        subprocess.run(["make", "-C", HERE], check=True, capture_output=True)


def _as_u8(data) -> np.ndarray:
    if isinstance(data, str):
        data = data.encode("latin-1")
    if isinstance(data, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(data), dtype=np.uint8)
    return np.ascontiguousarray(data, dtype=np.uint8)


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def pack_patterns(patterns):
    """list of bytes/str -> (bytes u8[total], offs u64[npat+1])."""
    arrs = [_as_u8(p) for p in patterns]
    offs = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        offs[1:] = np.cumsum([a.size for a in arrs], dtype=np.uint64)
    data = np.concatenate(arrs) if arrs and offs[-1] else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(data, dtype=np.uint8), offs


# ------------------------------------------------------------------------------------------
# liboracle.so
# ------------------------------------------------------------------------------------------
class _BitVec(C.Structure):
    _fields_ = [("nbits", C.c_uint64), ("nwords", C.c_uint64), ("nsuper", C.c_uint64),
                ("nsub", C.c_uint64), ("bits", _u64p), ("super_", _u32p), ("sub", _u16p),
                ("ones", C.c_uint64)]


class _Wavelet(C.Structure):
    _fields_ = [("n", C.c_uint64), ("lv", _BitVec * 8)]


class _Index(C.Structure):
    _fields_ = [("n", C.c_uint64), ("text", _u8p), ("bwt", _u8p), ("sa", _u32p),
                ("C", C.c_uint32 * 257), ("wt", _Wavelet), ("stride", C.c_uint32),
                ("ssa", _u32p), ("nsamp", C.c_uint64)]


_orc = None


def orc():
    global _orc
    if _orc is None:
        build()
        L = C.CDLL(ORACLE_SO)
        L.orc_sa_build.argtypes = [_u8p, C.c_uint64, _u32p]
        L.orc_sa_build.restype = C.c_int
        L.orc_sa_check.argtypes = [_u8p, C.c_uint64, _u32p]
        L.orc_sa_check.restype = C.c_int
        L.orc_bwt_from_sa.argtypes = [_u8p, C.c_uint64, _u32p, _u8p]
        L.orc_build_C.argtypes = [_u8p, C.c_uint64, _u32p]
        L.orc_bv_build.argtypes = [C.POINTER(_BitVec), _u8p, C.c_uint64]
        L.orc_bv_build_from_words.argtypes = [C.POINTER(_BitVec), _u64p, C.c_uint64, C.c_uint64]
        L.orc_bv_free.argtypes = [C.POINTER(_BitVec)]
        L.orc_bv_rank1.argtypes = [C.POINTER(_BitVec), C.c_uint64]
        L.orc_bv_rank1.restype = C.c_uint64
        L.orc_bv_get.argtypes = [C.POINTER(_BitVec), C.c_uint64]
        L.orc_bv_get.restype = C.c_uint8
        L.orc_wt_build.argtypes = [C.POINTER(_Wavelet), _u8p, C.c_uint64]
        L.orc_wt_free.argtypes = [C.POINTER(_Wavelet)]
        L.orc_wt_rank.argtypes = [C.POINTER(_Wavelet), C.c_uint8, C.c_uint64]
        L.orc_wt_rank.restype = C.c_uint64
        L.orc_wt_access.argtypes = [C.POINTER(_Wavelet), C.c_uint64]
        L.orc_wt_access.restype = C.c_uint8
        L.orc_index_build.argtypes = [_u8p, C.c_uint64, C.c_uint32]
        L.orc_index_build.restype = C.POINTER(_Index)
        L.orc_index_from_sa.argtypes = [_u8p, C.c_uint64, _u32p, C.c_uint32]
        L.orc_index_from_sa.restype = C.POINTER(_Index)
        L.orc_index_from_bwt.argtypes = [_u8p, C.c_uint64, _u32p, C.c_uint64, C.c_uint32]
        L.orc_index_from_bwt.restype = C.POINTER(_Index)
        L.orc_index_free.argtypes = [C.POINTER(_Index)]
        L.orc_count.argtypes = [C.POINTER(_Index), _u8p, C.c_uint64, _u64p, _u64p, _u64p]
        L.orc_count.restype = C.c_uint64
        L.orc_LF.argtypes = [C.POINTER(_Index), C.c_uint64]
        L.orc_LF.restype = C.c_uint64
        L.orc_locate.argtypes = [C.POINTER(_Index), _u8p, C.c_uint64, C.c_uint64, _u64p,
                                 C.c_uint64, _i32p, _u64p]
        L.orc_locate.restype = C.c_uint64
        L.orc_count_batch.argtypes = [C.POINTER(_Index), _u8p, _u64p, C.c_uint64, _u64p, _u64p,
                                      _u64p, C.c_int]
        L.orc_locate_batch.argtypes = [C.POINTER(_Index), _u8p, _u64p, C.c_uint64, C.c_uint64,
                                       _u64p, _u64p, C.c_uint64, _i32p, _u64p, C.c_int]
        L.orc_locate_batch.restype = C.c_uint64
        _orc = L
    return _orc


def sa_build(text) -> np.ndarray:
    t = _as_u8(text)
    sa = np.zeros(t.size, dtype=np.uint32)
    if orc().orc_sa_build(_ptr(t, _u8p), t.size, _ptr(sa, _u32p)) != 0:
        raise MemoryError("orc_sa_build")
    return sa


def sa_check(text, sa) -> int:
    t = _as_u8(text)
    sa = np.ascontiguousarray(sa, dtype=np.uint32)
    assert sa.size == t.size
    return int(orc().orc_sa_check(_ptr(t, _u8p), t.size, _ptr(sa, _u32p)))


class OracleBitVector:
    def __init__(self, bits01=None, words=None, nbits=None):
        self._bv = _BitVec()
        if words is not None:
            w = np.ascontiguousarray(words, dtype=np.uint64)
            orc().orc_bv_build_from_words(C.byref(self._bv), _ptr(w, _u64p), w.size, int(nbits))
        else:
            b = _as_u8(bits01)
            orc().orc_bv_build(C.byref(self._bv), _ptr(b, _u8p), b.size)

    def rank1(self, i): return int(orc().orc_bv_rank1(C.byref(self._bv), int(i)))
    def get(self, i): return int(orc().orc_bv_get(C.byref(self._bv), int(i)))
    def __len__(self): return int(self._bv.nbits)

    def __del__(self):
        try:
            orc().orc_bv_free(C.byref(self._bv))
        except Exception:
            pass


def _bv_arrays(bv: _BitVec):
    words = np.ctypeslib.as_array(bv.bits, shape=(bv.nwords,)).copy() if bv.nwords else np.zeros(0, np.uint64)
    sup = np.ctypeslib.as_array(bv.super_, shape=(bv.nsuper,)).copy() if bv.nsuper else np.zeros(0, np.uint32)
    sub = np.ctypeslib.as_array(bv.sub, shape=(bv.nsub,)).copy() if bv.nsub else np.zeros(0, np.uint16)
    return words, sup, sub


class OracleWavelet:
    def __init__(self, seq):
        s = _as_u8(seq)
        self._wt = _Wavelet()
        orc().orc_wt_build(C.byref(self._wt), _ptr(s, _u8p), s.size)

    def rank(self, c, i): return int(orc().orc_wt_rank(C.byref(self._wt), int(c), int(i)))
    def access(self, i): return int(orc().orc_wt_access(C.byref(self._wt), int(i)))
    def level_arrays(self, level): return _bv_arrays(self._wt.lv[level])

    def __del__(self):
        try:
            orc().orc_wt_free(C.byref(self._wt))
        except Exception:
            pass


class OracleIndex:
    """Restated cs::FMIndex (src/api/fm_index.cpp)."""

    def __init__(self, text=None, stride=32, sa=None, bwt=None, ssa=None):
        L = orc()
        self.stride = int(stride)
        if bwt is not None:
            b = _as_u8(bwt)
            s = np.ascontiguousarray(ssa if ssa is not None else np.zeros(0), dtype=np.uint32)
            self._p = L.orc_index_from_bwt(_ptr(b, _u8p), b.size, _ptr(s, _u32p), s.size, self.stride)
        else:
            t = _as_u8(text)
            if sa is None:
                self._p = L.orc_index_build(_ptr(t, _u8p), t.size, self.stride)
            else:
                s = np.ascontiguousarray(sa, dtype=np.uint32)
                self._p = L.orc_index_from_sa(_ptr(t, _u8p), t.size, _ptr(s, _u32p), self.stride)
        if not self._p:
            raise MemoryError("oracle index build failed")

    def __del__(self):
        try:
            if self._p:
                orc().orc_index_free(self._p)
        except Exception:
            pass

    @property
    def n(self): return int(self._p.contents.n)

    @property
    def sa(self):
        c = self._p.contents
        return np.ctypeslib.as_array(c.sa, shape=(c.n,)).copy() if c.n and c.sa else np.zeros(0, np.uint32)

    @property
    def bwt(self):
        c = self._p.contents
        return np.ctypeslib.as_array(c.bwt, shape=(c.n,)).copy() if c.n else np.zeros(0, np.uint8)

    @property
    def C(self): return np.array(self._p.contents.C[:], dtype=np.uint32)

    @property
    def ssa(self):
        c = self._p.contents
        return np.ctypeslib.as_array(c.ssa, shape=(c.nsamp,)).copy() if c.nsamp else np.zeros(0, np.uint32)

    def occ(self, c, i): return int(orc().orc_wt_rank(C.byref(self._p.contents.wt), int(c), int(i)))
    def LF(self, i): return int(orc().orc_LF(self._p, int(i)))

    def count_full(self, pattern):
        """-> (count, sp, ep, executed_steps)"""
        p = _as_u8(pattern)
        sp, ep, st = C.c_uint64(), C.c_uint64(), C.c_uint64()
        c = orc().orc_count(self._p, _ptr(p, _u8p), p.size, C.byref(sp), C.byref(ep), C.byref(st))
        return int(c), int(sp.value), int(ep.value), int(st.value)

    def count(self, pattern): return self.count_full(pattern)[0]

    def locate(self, pattern, limit=100000):
        """-> (positions list in SA-row order, status)"""
        p = _as_u8(pattern)
        cnt = self.count(p) if p.size else 0
        cap = max(1, min(cnt, int(limit)))
        out = np.zeros(cap, dtype=np.uint64)
        st = C.c_int32()
        k = orc().orc_locate(self._p, _ptr(p, _u8p), p.size, int(limit), _ptr(out, _u64p), cap,
                             C.byref(st), None)
        return out[:k].tolist(), int(st.value)

    def count_batch(self, data, offs, nthreads=0, want_steps=False):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        npat = offs.size - 1
        counts = np.zeros(npat, dtype=np.uint64)
        sp_ep = np.zeros(2 * npat, dtype=np.uint64)
        steps = np.zeros(npat, dtype=np.uint64)
        orc().orc_count_batch(self._p, _ptr(data, _u8p), _ptr(offs, _u64p), npat,
                              _ptr(counts, _u64p), _ptr(sp_ep, _u64p), _ptr(steps, _u64p),
                              int(nthreads))
        if want_steps:
            return counts, sp_ep.reshape(-1, 2), steps
        return counts, sp_ep.reshape(-1, 2)

    def locate_batch(self, data, offs, limit=100000, nthreads=0):
        """-> (out_offs u64[npat+1], out_pos u64[total], status i32[npat], lf_steps_total)"""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        npat = offs.size - 1
        out_offs = np.zeros(npat + 1, dtype=np.uint64)
        total = orc().orc_locate_batch(self._p, _ptr(data, _u8p), _ptr(offs, _u64p), npat,
                                       int(limit), _ptr(out_offs, _u64p), None, 0, None, None,
                                       int(nthreads))
        out_pos = np.zeros(max(1, total), dtype=np.uint64)
        status = np.zeros(npat, dtype=np.int32)
        lf = C.c_uint64()
        orc().orc_locate_batch(self._p, _ptr(data, _u8p), _ptr(offs, _u64p), npat, int(limit),
                               _ptr(out_offs, _u64p), _ptr(out_pos, _u64p), int(total),
                               _ptr(status, _i32p), C.byref(lf), int(nthreads))
        return out_offs, out_pos[:total], status, int(lf.value)


# ------------------------------------------------------------------------------------------
# _ref/libcsref.so (the compiled reference)
# ------------------------------------------------------------------------------------------
_ref = None


def ref_available() -> bool:
    if not os.path.exists(REF_SO):
        try:
            build()
        except Exception:
            return False
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        if not ref_available():
            raise RuntimeError("oracle/_ref/libcsref.so is not built (needs /root/reference)")
        L = C.CDLL(REF_SO)
        vp = C.c_void_p
        L.csref_last_error.restype = C.c_char_p
        L.csref_build_from_text.argtypes = [_u8p, C.c_uint64, C.c_uint32]
        L.csref_build_from_text.restype = vp
        L.csref_inject.argtypes = [_u8p, C.c_uint64, _u32p, C.c_uint32]
        L.csref_inject.restype = vp
        L.csref_inject_bwt.argtypes = [_u8p, C.c_uint64, _u32p, C.c_uint64, C.c_uint32]
        L.csref_inject_bwt.restype = vp
        L.csref_inject_planes.argtypes = [C.c_uint64, C.POINTER(_u64p), _u32p, _u8p, _u32p, C.c_uint64, C.c_uint32]
        L.csref_inject_planes.restype = vp
        L.csref_destroy.argtypes = [vp]
        L.csref_n.argtypes = [vp]
        L.csref_n.restype = C.c_uint64
        L.csref_count.argtypes = [vp, _u8p, C.c_uint64]
        L.csref_count.restype = C.c_uint64
        L.csref_locate.argtypes = [vp, _u8p, C.c_uint64, C.c_uint64, _u64p, C.c_uint64]
        L.csref_locate.restype = C.c_int64
        L.csref_count_many.argtypes = [vp, _u8p, _u64p, C.c_uint64, _u64p, C.c_int]
        L.csref_locate_many.argtypes = [vp, _u8p, _u64p, C.c_uint64, C.c_uint64, _i64p, _u64p, C.c_int]
        L.csref_locate_many.restype = C.c_uint64
        L.csref_extract.argtypes = [vp, C.c_uint64, C.c_uint64, _u8p]
        L.csref_extract.restype = C.c_uint64
        L.csref_get_sa.argtypes = [vp, _u32p]
        L.csref_get_bwt.argtypes = [vp, _u8p]
        L.csref_get_C.argtypes = [vp, _u32p]
        L.csref_ssa_size.argtypes = [vp]
        L.csref_ssa_size.restype = C.c_uint64
        L.csref_get_ssa.argtypes = [vp, _u32p]
        L.csref_occ.argtypes = [vp, C.c_uint8, C.c_uint64]
        L.csref_occ.restype = C.c_uint64
        L.csref_LF.argtypes = [vp, C.c_uint64]
        L.csref_LF.restype = C.c_uint64
        L.csref_build_sa_naive.argtypes = [_u8p, C.c_uint64, _u32p]
        L.csref_wt_build.argtypes = [_u8p, C.c_uint64]
        L.csref_wt_build.restype = vp
        L.csref_wt_destroy.argtypes = [vp]
        L.csref_wt_rank.argtypes = [vp, C.c_uint8, C.c_uint64]
        L.csref_wt_rank.restype = C.c_uint64
        L.csref_wt_access.argtypes = [vp, C.c_uint64]
        L.csref_wt_access.restype = C.c_uint8
        L.csref_wt_level_sizes.argtypes = [vp, C.c_int, _u64p, _u64p, _u64p]
        L.csref_wt_level_sizes.restype = C.c_uint64
        L.csref_wt_level_dump.argtypes = [vp, C.c_int, _u64p, _u32p, _u16p]
        L.csref_bv_build.argtypes = [_u8p, C.c_uint64]
        L.csref_bv_build.restype = vp
        L.csref_bv_build_from_words.argtypes = [_u64p, C.c_uint64, C.c_uint64]
        L.csref_bv_build_from_words.restype = vp
        L.csref_bv_destroy.argtypes = [vp]
        L.csref_bv_rank1.argtypes = [vp, C.c_uint64]
        L.csref_bv_rank1.restype = C.c_uint64
        L.csref_bv_get.argtypes = [vp, C.c_uint64]
        L.csref_bv_get.restype = C.c_uint8
        L.csref_bv_size.argtypes = [vp]
        L.csref_bv_size.restype = C.c_uint64
        L.csref_hardware_threads.restype = C.c_int
        _ref = L
    return _ref


def ref_sa_naive(text) -> np.ndarray:
    t = _as_u8(text)
    sa = np.zeros(t.size, dtype=np.uint32)
    ref().csref_build_sa_naive(_ptr(t, _u8p), t.size, _ptr(sa, _u32p))
    return sa


class RefBitVector:
    def __init__(self, bits01=None, words=None, nbits=None):
        if words is not None:
            w = np.ascontiguousarray(words, dtype=np.uint64)
            self._h = ref().csref_bv_build_from_words(_ptr(w, _u64p), w.size, int(nbits))
        else:
            b = _as_u8(bits01)
            self._h = ref().csref_bv_build(_ptr(b, _u8p), b.size)

    def rank1(self, i): return int(ref().csref_bv_rank1(self._h, int(i)))
    def get(self, i): return int(ref().csref_bv_get(self._h, int(i)))
    def __len__(self): return int(ref().csref_bv_size(self._h))

    def __del__(self):
        try:
            ref().csref_bv_destroy(self._h)
        except Exception:
            pass


class RefWavelet:
    def __init__(self, seq):
        s = _as_u8(seq)
        self._h = ref().csref_wt_build(_ptr(s, _u8p), s.size)

    def rank(self, c, i): return int(ref().csref_wt_rank(self._h, int(c), int(i)))
    def access(self, i): return int(ref().csref_wt_access(self._h, int(i)))

    def level_arrays(self, level):
        nw, ns, nb = C.c_uint64(), C.c_uint64(), C.c_uint64()
        ref().csref_wt_level_sizes(self._h, level, C.byref(nw), C.byref(ns), C.byref(nb))
        words = np.zeros(nw.value, np.uint64)
        sup = np.zeros(ns.value, np.uint32)
        sub = np.zeros(nb.value, np.uint16)
        ref().csref_wt_level_dump(self._h, level, _ptr(words, _u64p), _ptr(sup, _u32p), _ptr(sub, _u16p))
        return words, sup, sub

    def __del__(self):
        try:
            ref().csref_wt_destroy(self._h)
        except Exception:
            pass


class RefIndex:
    """The compiled, unmodified cs::FMIndex."""

    def __init__(self, text=None, stride=32, sa=None, bwt=None, ssa=None, planes=None, C_array=None, n=None):
        L = ref()
        self.stride = int(stride)
        if planes is not None:
            # eight packed bit planes (u64, LSB-first) -> cs::BitVector::build_from_words
            nwords = (int(n) + 63) // 64
            keep = [np.ascontiguousarray(p[:nwords], dtype=np.uint64) for p in planes]
            assert len(keep) == 8 and all(k.size == nwords for k in keep)
            arr = (_u64p * 8)(*[_ptr(k, _u64p) for k in keep])
            cc = np.ascontiguousarray(C_array, dtype=np.uint32)
            b = _as_u8(bwt) if bwt is not None else None
            s = np.ascontiguousarray(ssa, dtype=np.uint32) if ssa is not None else None
            self._h = L.csref_inject_planes(int(n), arr, _ptr(cc, _u32p), _ptr(b, _u8p) if b is not None else None,
                                            _ptr(s, _u32p) if s is not None else None,
                                            s.size if s is not None else 0, self.stride)
        elif bwt is not None:
            b = _as_u8(bwt)
            s = np.ascontiguousarray(ssa if ssa is not None else np.zeros(0), dtype=np.uint32)
            self._h = L.csref_inject_bwt(_ptr(b, _u8p), b.size, _ptr(s, _u32p) if s.size else None,
                                         s.size, self.stride)
        else:
            t = _as_u8(text)
            if sa is None:
                self._h = L.csref_build_from_text(_ptr(t, _u8p), t.size, self.stride)
            else:
                s = np.ascontiguousarray(sa, dtype=np.uint32)
                self._h = L.csref_inject(_ptr(t, _u8p), t.size, _ptr(s, _u32p), self.stride)

    def __del__(self):
        try:
            ref().csref_destroy(self._h)
        except Exception:
            pass

    @property
    def n(self): return int(ref().csref_n(self._h))

    @property
    def sa(self):
        out = np.zeros(self.n, np.uint32)
        ref().csref_get_sa(self._h, _ptr(out, _u32p))
        return out

    @property
    def bwt(self):
        out = np.zeros(self.n, np.uint8)
        ref().csref_get_bwt(self._h, _ptr(out, _u8p))
        return out

    @property
    def C(self):
        out = np.zeros(257, np.uint32)
        ref().csref_get_C(self._h, _ptr(out, _u32p))
        return out

    @property
    def ssa(self):
        out = np.zeros(ref().csref_ssa_size(self._h), np.uint32)
        ref().csref_get_ssa(self._h, _ptr(out, _u32p))
        return out

    def occ(self, c, i): return int(ref().csref_occ(self._h, int(c), int(i)))
    def LF(self, i): return int(ref().csref_LF(self._h, int(i)))

    def count(self, pattern):
        p = _as_u8(pattern)
        return int(ref().csref_count(self._h, _ptr(p, _u8p), p.size))

    def locate(self, pattern, limit=100000):
        """-> (positions list in SA-row order, status) ; status 1 == the reference threw"""
        p = _as_u8(pattern)
        cnt = self.count(p) if p.size else 0
        cap = max(1, min(cnt, int(limit)))
        out = np.zeros(cap, dtype=np.uint64)
        k = ref().csref_locate(self._h, _ptr(p, _u8p), p.size, int(limit), _ptr(out, _u64p), cap)
        if k < 0:
            msg = ref().csref_last_error().decode()
            return [], (ORC_LF_WALK_EXCEEDED if "LF walk" in msg else ORC_SSA_OOB)
        return out[:k].tolist(), ORC_OK

    def count_batch(self, data, offs, nthreads=1):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        npat = offs.size - 1
        out = np.zeros(npat, dtype=np.uint64)
        ref().csref_count_many(self._h, _ptr(data, _u8p), _ptr(offs, _u64p), npat, _ptr(out, _u64p),
                               int(nthreads))
        return out

    def locate_batch(self, data, offs, limit=100000, nthreads=1, keep_positions=False):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        npat = offs.size - 1
        out_n = np.zeros(npat, dtype=np.int64)
        pos = np.zeros(npat * int(limit), dtype=np.uint64) if keep_positions else None
        total = ref().csref_locate_many(self._h, _ptr(data, _u8p), _ptr(offs, _u64p), npat, int(limit),
                                        _ptr(out_n, _i64p), _ptr(pos, _u64p) if keep_positions else None,
                                        int(nthreads))
        return int(total), out_n, pos


def hardware_threads() -> int:
    return os.cpu_count() or 1
