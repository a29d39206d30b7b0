"""B200-native FM-index query engine (count / locate hot path of cs::FMIndex).

csrc/    CUDA kernels for sm_100a + the C ABI (include/csfm.h) -> libcsfm.so
host/    C++ drop-in mirror of the reference header (src/api/fm_index.hpp)
binding  ctypes access to the same C ABI for the parity tests and bench.py
"""
from .binding import (BuildParams, CsfmError, FMIndex, LIB_PATH, SIGNATURES, host_alloc, host_free, lib,  # noqa: F401
                      pack_patterns, Q_OK, Q_LF_WALK_EXCEEDED, Q_SSA_OOB, BUILD_DEFAULT, BUILD_NO_COMPACT,
                      BUILD_KEEP_SA, BUILD_LAYOUT_BINARY64, BUILD_NO_KMER_TABLE, BUILD_NO_TEXT_CHECK, BUILD_FORCE_TEXT_CHECK, BUILD_LARGE_TABLE, BUILD_LAYOUT_NIBBLE128, BUILD_ROW_SAMPLES, LF_WALK_MESSAGE)
