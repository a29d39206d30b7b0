#!/usr/bin/env bash
# Builds libcs_b200.so: the C++ drop-in for the reference's static library `cs`
# (cs::FMIndex over the C ABI of ../libcsfm.so). Plain g++, no CUDA headers needed.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/libcs_b200.so"
CXX=/usr/bin/g++
[[ -x "$CXX" ]] || CXX=g++
TOOL="$HERE/cs_benchmark_batch"
if [[ -f "$OUT" && -f "$TOOL" && "$TOOL" -nt "$HERE/tools/benchmark_batch.cpp" && "$TOOL" -nt "$OUT" && "$OUT" -nt "$HERE/src/api/fm_index.cpp" && "$OUT" -nt "$HERE/src/serialization/csidx.cpp" && "$OUT" -nt "$HERE/src/serialization/csidx.hpp" && "$OUT" -nt "$HERE/src/api/fm_index.hpp" && "$OUT" -nt "$HERE/../libcsfm.so" && "${1:-}" != "-f" ]]; then
  exit 0
fi
"$CXX" -std=c++20 -O2 -fPIC -Wall -Wextra -shared -o "$OUT" "$HERE/src/api/fm_index.cpp" "$HERE/src/serialization/csidx.cpp" \
  -L"$HERE/.." -lcsfm -Wl,-rpath,'$ORIGIN/..'
# the --batch companion of the reference's tools/benchmark.cpp, on the drop-in class
"$CXX" -std=c++20 -O2 -Wall -Wextra -o "$TOOL" "$HERE/tools/benchmark_batch.cpp" -L"$HERE" -lcs_b200 -L"$HERE/.." -lcsfm \
  -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/..'
