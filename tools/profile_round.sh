#!/usr/bin/env bash
# Round-end evidence on a B200 (run through gpurun): the plain bench first, then the launch list and
# one `ncu --set full` capture per kernel of interest. Launch order of count2_kernel in
# `bench.py --steps 3 --warmup 3`: 8 instrumented + 3 warm-up + 3 timed (default), 7 stepping-only,
# 1 instrumented + 3 warm-up + 3 timed on the large-table index.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 200 $B > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'count2|walk2|widen|narrow|expand_rows|rows_to|DeviceScan|kmer' -c 600 --csv --log-file gpurun_out/launches_bench_c3.csv $B > gpurun_out/prof_ll.json 2> gpurun_out/prof_ll.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:count2_kernel -s 12 -c 1 -f -o gpurun_out/count2_default $B --no-locate > /dev/null 2> gpurun_out/prof_a.err
if [[ "${1:-}" == "all" ]]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:count2_kernel -s 26 -c 1 -f -o gpurun_out/count2_large $B --no-locate > /dev/null 2> gpurun_out/prof_b.err
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk2_kernel -s 3 -c 1 -f -o gpurun_out/walk2 $B --no-large-table > /dev/null 2> gpurun_out/prof_c.err
fi
ls -la gpurun_out/*.ncu-rep
