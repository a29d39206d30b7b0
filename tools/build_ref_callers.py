#!/usr/bin/env python
"""tools/build_ref_callers.py — compile the reference's OWN callers, unmodified, against the
drop-in header (host/src/api/fm_index.hpp) and the drop-in library (libcs_b200.so).

Runs only where /root/reference exists (the build container). The reference sources are staged
into build/ref_callers/stage/ just long enough to compile them — their `#include
"../src/api/fm_index.hpp"` then resolves to OUR header through a symlink — and are removed again;
only the binaries stay (build/ is git-ignored but travels to the GPU box with gpurun):

    build/ref_callers/bin/{cs_query, cs_bench, benchmark, build_index, fm_search_tests, cs_tests, debug_fm}

With --expected it also builds the same callers against the REAL reference sources, runs them here
on the CPU and stores their stable output lines in tests/golden/ref_callers_expected.json, which
tests/test_gpu_dropin.py compares with the output of the drop-in binaries on the GPU.
"""
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PKG = os.path.join(ROOT, "compressed-fm-index-implementation-with-learned-optimizations_b200")
OUT = os.path.join(ROOT, "build", "ref_callers")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

CALLERS = {  # binary name -> reference source (CMakeLists.txt:74-138 names)
    "cs_query": "tools/query_cli.cpp", "cs_bench": "tools/bench.cpp", "benchmark": "tools/benchmark.cpp",
    "build_index": "tools/build_index.cpp", "fm_search_tests": "tests/fm_search_tests.cpp",
    "cs_tests": "tests/simple_tests.cpp", "debug_fm": "tests/debug_fm.cpp",
}


def build_dropin():
    stage = os.path.join(OUT, "stage")
    shutil.rmtree(stage, ignore_errors=True)
    os.makedirs(os.path.join(OUT, "bin"), exist_ok=True)
    try:
        for sub in ("tools", "tests"):
            os.makedirs(os.path.join(stage, sub))
        os.symlink(os.path.join(PKG, "host", "src"), os.path.join(stage, "src"))
        for name, src in CALLERS.items():
            dst = os.path.join(stage, src)
            shutil.copyfile(os.path.join(REF, src), dst)
            # asserts stay live (no -DNDEBUG): the reference's tests are assert-based
            subprocess.run([CXX, "-std=c++20", "-O2", "-o", os.path.join(OUT, "bin", name), dst,
                            "-L" + os.path.join(PKG, "host"), "-lcs_b200", "-L" + PKG, "-lcsfm",
                            "-Wl,-rpath," + os.path.join(PKG, "host"), "-Wl,-rpath," + PKG,
                            "-Wl,-rpath,$ORIGIN/../../../" + os.path.basename(PKG) + "/host",
                            "-Wl,-rpath,$ORIGIN/../../../" + os.path.basename(PKG)], check=True)
    finally:
        shutil.rmtree(stage, ignore_errors=True)
    print("drop-in callers:", sorted(os.listdir(os.path.join(OUT, "bin"))))


def stable_lines(name, text):
    """Lines that do not depend on timing."""
    keep = []
    for line in text.splitlines():
        if name == "benchmark":
            if re.search(r"Total matches|Text size|Queries:|Pattern len", line):
                keep.append(line.strip())
        elif re.search(r"\bms\b|QPS|Time|time|μs|us\b", line) and name in ("build_index",):
            continue
        else:
            keep.append(line.rstrip())
    return keep


def build_expected():
    refbin = os.path.join(OUT, "refbin")
    os.makedirs(refbin, exist_ok=True)
    tus = [os.path.join(REF, p) for p in ("src/api/fm_index.cpp", "src/core/wavelet.cpp", "src/core/bitvector.cpp")]
    for name, src in CALLERS.items():
        subprocess.run([CXX, "-std=c++20", "-O3", "-mavx2", "-mbmi2", "-mpopcnt", "-DCS_AVX2", "-I" + REF + "/src",
                        "-I" + REF + "/include", "-o", os.path.join(refbin, name), os.path.join(REF, src)] + tus, check=True)
    runs = {
        "cs_query sample banana": ("cs_query", ["{sample}", "banana"]),
        "cs_query sample ana": ("cs_query", ["{sample}", "ana"]),
        "cs_query sample band": ("cs_query", ["{sample}", "band"]),
        "cs_query example quick": ("cs_query", ["{example}", "quick"]),
        "cs_query example algorithm": ("cs_query", ["{example}", "algorithm"]),
        "cs_query example the": ("cs_query", ["{example}", "the"]),
        "cs_bench sample": ("cs_bench", ["{sample}"]),
        "cs_bench example": ("cs_bench", ["{example}"]),
        "debug_fm": ("debug_fm", []),
        "fm_search_tests": ("fm_search_tests", []),
        "cs_tests": ("cs_tests", []),
        "benchmark": ("benchmark", []),
    }
    paths = {"sample": os.path.join(REF, "sample.txt"), "example": os.path.join(REF, "example.txt")}
    expected = {}
    for key, (exe, args) in runs.items():
        argv = [os.path.join(refbin, exe)] + [a.format(**paths) for a in args]
        print("running reference", key, "...", flush=True)
        r = subprocess.run(argv, capture_output=True, text=True, errors="replace")
        err = [l for l in r.stderr.splitlines() if not l.startswith("[TIMER]")]
        expected[key] = {"exe": exe, "args": args, "returncode_is_zero": r.returncode == 0,
                         "stdout": stable_lines(exe, r.stdout),
                         "stderr": [re.sub(r"^.*?(Assertion)", r"\1", l) for l in err]}
    json.dump({"generator": "tools/build_ref_callers.py --expected", "runs": expected},
              open(os.path.join(ROOT, "tests", "golden", "ref_callers_expected.json"), "w"), indent=1)
    shutil.rmtree(refbin, ignore_errors=True)


if __name__ == "__main__":
    if not os.path.exists(os.path.join(REF, "tools", "benchmark.cpp")):
        print("reference sources not present: keeping prebuilt build/ref_callers/bin if any")
        sys.exit(0)
    subprocess.run(["bash", os.path.join(PKG, "build.sh")], check=True)
    subprocess.run(["bash", os.path.join(PKG, "host", "build.sh")], check=True)
    build_dropin()
    if "--expected" in sys.argv:
        build_expected()
