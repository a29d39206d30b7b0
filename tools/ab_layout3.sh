# A/B runs of the layout-3 kernels (round 2, profiles/README.md §R2.2). The experiment libraries are built beside the
# product library from the same sources (they are git-ignored):
#   PK=compressed-fm-index-implementation-with-learned-optimizations_b200
#   CSFM_OUT=$PK/libcsfm_vA.so CSFM_NVCC_EXTRA="-DCSFM_DNA_SYMBOL_BRANCHY -DCSFM_WALK3_CTAS=8 -DCSFM_COUNT3_CTAS=8" bash $PK/build.sh -f
#   CSFM_OUT=$PK/libcsfm_vB.so CSFM_NVCC_EXTRA="-DCSFM_DNA_SYMBOL_BRANCHY" bash $PK/build.sh -f      (6 CTAs/SM, the default now)
#   CSFM_OUT=$PK/libcsfm_vC.so bash $PK/build.sh -f                                                  (+ branch-free symbol pick)
# and selected through CSFM_LIB. Result on one B200: C2 count 7.06 / 7.38 / 7.38e9 q/s, C4 locate 2.97 / 3.00 / 3.01e9 occ/s.
mkdir -p gpurun_out
PK=compressed-fm-index-implementation-with-learned-optimizations_b200
for v in vA vB vC main; do
  if [ $v = main ]; then unset CSFM_LIB; else export CSFM_LIB=$PWD/$PK/libcsfm_$v.so; fi
  timeout 300 python bench.py --workload c2 --no-configs --steps 100 --no-cpu-baseline --locate-patterns 300000 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || echo "fail $v"
done
python - <<'PY'
import json
for v in ("vA","vB","vC","main"):
    try:
        d=json.load(open(f"gpurun_out/ab_{v}.json"))
        print(v, "c2 count %.3e"%d["value"], "locate %.3e"%d["locate"]["value"])
    except Exception as e: print(v, "ERR", e)
PY
