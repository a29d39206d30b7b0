#!/usr/bin/env python3
"""Stand-in for compute-sanitizer (closed on the GPU pool): builds libcsfm_check.so with -DCSFM_BOUNDS_CHECK — every
data-dependent address of the query kernels is range-checked before the load and a violation traps — runs the parity
suites against it (CSFM_LIB), and proves in a subprocess that the check fires on a corrupted index line.

    python tools/bounds_check.py [pytest args]        (on a GPU box; writes gpurun_out/bounds_check.log)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "compressed-fm-index-implementation-with-learned-optimizations_b200")
LIB = os.path.join(PKG, "libcsfm_check.so")

SELFTEST = r'''
import numpy as np, sys
sys.path.insert(0, %r)
import csfm_b200 as fm
rng = np.random.default_rng(1)
text = np.concatenate([rng.integers(1, 60, 50000).astype(np.uint8), np.zeros(1, np.uint8)])
idx = fm.FMIndex.build_from_text(text, fm.BuildParams(ssa_stride=8), flags=fm.BUILD_NO_KMER_TABLE)
blob = idx.blob_to_host()
nblk, off_levels = (int(x) for x in np.frombuffer(blob[40:56].tobytes(), np.uint64))
lines = blob[off_levels: off_levels + nblk * 128].reshape(-1, 128)
lines[:, 0:16] = 0xF0                             # the first four counters of every level-0 line: ranks far beyond n
bad = fm.FMIndex.from_host_blob(blob)
d, o = fm.pack_patterns([text[100:110].tobytes()] * 64)
print("counting on the corrupted index ...", flush=True)
bad.count_batch(d, o)
print("NO TRAP")
''' % ROOT


def main():
    env = dict(os.environ, CSFM_OUT=LIB, CSFM_NVCC_EXTRA="-DCSFM_BOUNDS_CHECK")
    subprocess.run(["bash", os.path.join(PKG, "build.sh"), "-f"], check=True, env=env)
    env = dict(os.environ, CSFM_LIB=LIB)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "bounds_check.log"), "w")
    r = subprocess.run([sys.executable, "-c", SELFTEST], env=env, capture_output=True, text=True)
    fired = "CSFM_BOUNDS_CHECK" in (r.stdout + r.stderr) and "NO TRAP" not in r.stdout
    log.write(f"self-test (corrupted counters must trap): fired={fired} rc={r.returncode}\n")
    log.write("\n".join((r.stdout + r.stderr).splitlines()[:6]) + "\n\n")
    args = sys.argv[1:] or ["tests/test_gpu_parity.py", "tests/test_gpu_property.py", "tests/test_gpu_sa_builder.py", "-x", "-q", "-m", "gpu"]
    t = subprocess.run([sys.executable, "-m", "pytest"] + args, env=env, cwd=ROOT, capture_output=True, text=True)
    log.write("pytest " + " ".join(args) + f" with CSFM_LIB={os.path.relpath(LIB, ROOT)}\n")
    log.write("\n".join(t.stdout.splitlines()[-12:]) + "\n")
    log.close()
    print(open(log.name).read())
    sys.exit(0 if (fired and t.returncode == 0) else 1)


if __name__ == "__main__":
    main()
