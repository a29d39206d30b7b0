# A/B of the layout-3 count kernel on C2 (L2-resident): one lane per query (default while the index fits the L2)
# against the two-lane sub-warp form (CSFM_COUNT3_LANES=2)
mkdir -p gpurun_out
B="python bench.py --workload c2 --no-configs --no-locate --steps 100 --no-cpu-baseline"
$B > gpurun_out/l_1.json 2>/dev/null
CSFM_COUNT3_LANES=2 $B > gpurun_out/l_2.json 2>/dev/null
python - <<'PY'
import json
for f in ("l_1", "l_2"):
    d = json.load(open("gpurun_out/%s.json" % f)); r = d["roofline"]
    print(f, "%.3e q/s" % d["value"], "kernel ms min/p50 %.4f %.4f" % (r["kernel_ms_min"], r["kernel_ms_p50"]), "frac", round(r["frac"], 3), d["checks"])
PY
