mkdir -p gpurun_out
B="python bench.py --workload c2 --no-configs --no-locate --steps 100 --no-cpu-baseline"
$B > gpurun_out/k_def.json 2>/dev/null
CSFM_KMER_BUDGET_MB=40 $B > gpurun_out/k_40.json 2>/dev/null
CSFM_KMER_BUDGET_MB=140 $B > gpurun_out/k_140.json 2>/dev/null
python - <<'PY'
import json
for f in ("k_def","k_40","k_140"):
    d=json.load(open("gpurun_out/%s.json"%f)); print(f, "%.3e"%d["value"], d["roofline"]["kmer_k"], d["config"]["index_bytes"])
PY
