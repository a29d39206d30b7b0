# A/B of the layout-3 walk kernel on C4 (2^28 DNA, 123 MB of lines + samples: half of it hits the L2): the two-lane
# sub-warp form (default beyond the L2) against one lane per row (CSFM_WALK3_LANES=1)
mkdir -p gpurun_out
B="python bench.py --workload c2 --no-configs --steps 3 --warmup 3 --no-cpu-baseline"
CSFM_WALK3_LANES=2 $B > gpurun_out/w_2.json 2>/dev/null
CSFM_WALK3_LANES=1 $B > gpurun_out/w_1.json 2>/dev/null
python - <<'PY'
import json
for f in ("w_2", "w_1"):
    d = json.load(open("gpurun_out/%s.json" % f))["locate"]
    print(f, "%.3e occ/s" % d["value"], {k: d[k] for k in ("ms_per_batch", "positions_ok") if k in d}, d.get("roofline", {}).get("frac"))
PY
