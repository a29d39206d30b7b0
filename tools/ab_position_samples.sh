# A/B of the marked line form of layout 3 (suffix array sampled by text position, csrc/csfm_dna.cuh) against the
# reference's row sampling (CSFM_NO_POSITION_SAMPLES=1, the 192-row line): C2 count (does the wider line cost the search
# anything?) and C4 locate at 1 M patterns, the marked walk with one lane and with two lanes per row.
mkdir -p gpurun_out
B="python bench.py --workload c2 --no-configs --steps 100 --no-cpu-baseline"
CSFM_NO_POSITION_SAMPLES=1 $B > gpurun_out/p_rows.json 2> gpurun_out/p_rows.err
CSFM_WALK3_LANES=1 $B > gpurun_out/p_m1.json 2> gpurun_out/p_m1.err
CSFM_WALK3_LANES=2 $B > gpurun_out/p_m2.json 2> gpurun_out/p_m2.err
python - <<'PY'
import json
for f in ("p_rows", "p_m1", "p_m2"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f)); l = d["locate"]
        print(f, "c2 count %.3e" % d["value"], "| c4 locate %.3e occ/s" % l["value"], "steps/occ %.2f" % l["lf_steps_per_occurrence"],
              "index %d MB walk set %d MB" % (l["config"]["index_bytes"] >> 20, l["config"]["walk_working_set_bytes"] >> 20),
              l["checks"], l.get("walk_lengths", {}).get("max"), l.get("walk_lengths", {}).get("sum_equals_lf_steps_counted"),
              "e2e %.3e" % l["e2e"]["value"], l["resident_sa"].get("positions_equal_walk"))
    except Exception as e: print(f, "ERR", repr(e))
PY
