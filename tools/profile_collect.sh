#!/usr/bin/env bash
# After `tools/profile_round.sh <R> all` has run on the GPU box and its files are back in gpurun_out/: summarise the
# captures into profiles/ (tracked), write profiles/count_kernel_traffic.json stamped with the kernel-source hash, copy
# the launch list, regenerate the SASS excerpts.   usage: tools/profile_collect.sh <R>
set -euo pipefail
R=${1:?round prefix of the gpurun_out files, e.g. r2i}
cd "$(dirname "$0")/.."
OCC=$(python -c "import json;print(json.load(open('gpurun_out/${R}_prof_plain.json'))['locate']['occurrences_per_batch'])")
python tools/ncu_traffic.py c3=gpurun_out/${R}_count2_c3_default.ncu-rep:count2_kernel c3_stepping=gpurun_out/${R}_count2_c3_stepping.ncu-rep:count2_kernel \
  c3_large_table=gpurun_out/${R}_count2_c3_large.ncu-rep:count2_kernel c2=gpurun_out/${R}_count3_c2.ncu-rep:count3_kernel \
  c5=gpurun_out/${R}_count3_c5.ncu-rep:count3_kernel c4_walk=gpurun_out/${R}_walk3_c4.ncu-rep:walk3_kernel:$OCC \
  tma=gpurun_out/${R}_count2_tma.ncu-rep:count2_tma_kernel > /dev/null
for f in count2_c3_default count2_c3_stepping count3_c2 count3_c5 walk3_c4 count2_c3_large count2_tma; do
  python tools/ncu_summary.py gpurun_out/${R}_$f.ncu-rep > profiles/r2_${f}_ncu_full.json
done
cp gpurun_out/${R}_launches_bench_default.csv profiles/r2_launches_bench_default.csv
cp gpurun_out/${R}_prof_plain.json profiles/r2_bench_steps3_before_ncu.json
python tools/sass_excerpts.py r2 > /dev/null
python - <<'PY'
import json, bench
d = json.load(open("profiles/count_kernel_traffic.json"))
assert d["source_sha16"] == bench.kernel_sources_sha16()
print("sources", d["source_sha16"])
for k, v in d["captures"].items():
    print(f"{k:16s} {v['dram_bytes_per_launch'] / 1e6:9.1f} MB  {v.get('gpu__time_duration.sum')}")
PY
