# A/B runs of the C3 default count kernel: CTAs per SM (experiment build:
#   CSFM_OUT=$PK/libcsfm_v7.so CSFM_NVCC_EXTRA="-DCSFM_SHORTCUT_CTAS=7" bash $PK/build.sh -f) and the refill knobs (environment).
# Result on one B200 (profiles/README.md §R2.3): default 6.42e9, 7 CTAs 5.35e9, knobs within 1 %.
mkdir -p gpurun_out
PK=compressed-fm-index-implementation-with-learned-optimizations_b200
B="python bench.py --steps 100 --no-cpu-baseline --no-configs --no-locate --no-large-table"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/abc_$name.json 2> gpurun_out/abc_$name.err || echo "fail $name"; }
run main CSFM_X=0
[ -f $PK/libcsfm_v7.so ] && run ctas7 CSFM_LIB=$PWD/$PK/libcsfm_v7.so
run min6 CSFM_REFILL_MIN=6
run min4 CSFM_REFILL_MIN=4
run wait3 CSFM_REFILL_WAIT=3
run wait8 CSFM_REFILL_WAIT=8
run min6wait3 CSFM_REFILL_MIN=6 CSFM_REFILL_WAIT=3
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/abc_*.json")):
    try:
        d=json.load(open(f)); print(f.split("abc_")[1][:-5], "%.3e q/s"%d["value"], "min %.4f p50 %.4f ms"%(d["roofline"]["kernel_ms_min"], d["roofline"]["kernel_ms_p50"]))
    except Exception as e: print(f, "ERR", e)
PY
