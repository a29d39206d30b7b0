# A/B of L2 residency hints in the layout-3 kernels (CSFM_L2_HINTS, csrc/csfm_common.cuh): 0 none, 1 suffix-array samples
# and result stores evict-first, 2 + level lines and k-mer table entries evict-last. C2 count (56 MB of lines + table
# against 25 MB of patterns and counts streaming through per batch) and C4 locate (123 MB of lines + samples, 2 GB of
# positions written per batch). Experiment libraries (git-ignored), built beside the product library:
#   PK=compressed-fm-index-implementation-with-learned-optimizations_b200
#   for h in 0 1 2; do CSFM_OUT=$PK/libcsfm_h$h.so CSFM_NVCC_EXTRA="-DCSFM_L2_HINTS=$h" bash $PK/build.sh -f; done
mkdir -p gpurun_out
PK=compressed-fm-index-implementation-with-learned-optimizations_b200
for v in ${HINT_VARIANTS:-h0 h1 h2}; do
  export CSFM_LIB=$PWD/$PK/libcsfm_$v.so
  timeout 300 python bench.py --workload c2 --no-configs --steps 100 --no-cpu-baseline > gpurun_out/abh_$v.json 2> gpurun_out/abh_$v.err || echo "fail $v"
done
python - <<'PY'
import json, os
for v in os.environ.get("HINT_VARIANTS", "h0 h1 h2").split():
    try:
        d = json.load(open(f"gpurun_out/abh_{v}.json"))
        print(v, "c2 count %.3e" % d["value"], "c4 locate (1M patterns) %.3e" % d["locate"]["value"], d["locate"]["checks"])
    except Exception as e: print(v, "ERR", e)
PY
