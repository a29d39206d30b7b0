# A/B of the one-lane layout-3 kernels at 3 / 4 (main) / 5 / 6 CTAs per SM (C2 count, C4 locate), and of the count
# kernel's two forms on C5 (4e9 DNA, out of HBM). Experiment libraries (git-ignored), built beside the product library:
#   PK=compressed-fm-index-implementation-with-learned-optimizations_b200
#   for c in 3 5 6; do CSFM_OUT=$PK/libcsfm_t$c.so CSFM_NVCC_EXTRA="-DCSFM_WALK3T_CTAS=$c -DCSFM_COUNT3T_CTAS=$c" bash $PK/build.sh -f; done
mkdir -p gpurun_out
PK=compressed-fm-index-implementation-with-learned-optimizations_b200
export CSFM_WALK3_LANES=1 CSFM_COUNT3_LANES=1
for v in t3 main t5 t6; do
  if [ $v = main ]; then unset CSFM_LIB; else export CSFM_LIB=$PWD/$PK/libcsfm_$v.so; fi
  timeout 300 python bench.py --workload c2 --no-configs --steps 100 --no-cpu-baseline --locate-patterns 300000 > gpurun_out/abt_$v.json 2> gpurun_out/abt_$v.err || echo "fail $v"
done
unset CSFM_LIB CSFM_WALK3_LANES
for l in 2 1; do
  CSFM_COUNT3_LANES=$l timeout 300 python bench.py --workload c5 --no-configs --no-locate --steps 50 --no-cpu-baseline > gpurun_out/abt_c5_l$l.json 2> gpurun_out/abt_c5_l$l.err || echo "fail c5 $l"
done
python - <<'PY'
import json
for v in ("t3", "main", "t5", "t6"):
    try:
        d = json.load(open(f"gpurun_out/abt_{v}.json"))
        print(v, "c2 count %.3e" % d["value"], "c4 locate (300k patterns) %.3e" % d["locate"]["value"])
    except Exception as e: print(v, "ERR", e)
for l in (2, 1):
    try:
        d = json.load(open(f"gpurun_out/abt_c5_l{l}.json"))
        print("c5 lanes", l, "%.3e q/s" % d["value"], d["checks"])
    except Exception as e: print("c5", l, "ERR", e)
PY
