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
