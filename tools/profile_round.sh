#!/usr/bin/env bash
# Round-end evidence on a B200 (run through gpurun): the plain bench first (must exit 0 without ncu), then the launch
# list of the same command and one `ncu --set full` capture per kernel of interest. Launch order of the count kernel in
# a `bench.py --steps 3 --warmup 3` count leg: 8 instrumented launches, 3 warm-up, 3 timed; then (c3 only) the
# stepping-only leg: 1 instrumented, 3 warm-up, 3 timed. ncu's -k matches the function name without template
# arguments, so template variants are told apart by their position (-s).
set -x
mkdir -p gpurun_out
R=${1:-r2}
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 300 $B --locate-patterns 200000 > gpurun_out/${R}_prof_plain.json 2> gpurun_out/${R}_prof_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'count[23]|walk[23]|single|widen|narrow|expand_rows|rows_to|DeviceScan|kmer|dna_|nib_' -c 900 --csv \
  --log-file gpurun_out/${R}_launches_bench_default.csv $B --locate-patterns 200000 > gpurun_out/${R}_prof_ll.json 2> gpurun_out/${R}_prof_ll.err
# C3: default kernel (count2_kernel<true,false>, a timed launch) and the stepping kernel (count2_kernel<false,false>)
timeout 300 $NCU -k regex:'count2_kernel' -s 12 -o gpurun_out/${R}_count2_c3_default $B --no-locate --no-configs --no-large-table > /dev/null 2> gpurun_out/${R}_prof_a.err
timeout 300 $NCU -k regex:'count2_kernel' -s 17 -o gpurun_out/${R}_count2_c3_stepping $B --no-locate --no-configs --no-large-table > /dev/null 2> gpurun_out/${R}_prof_b.err
# C2 / C5: layout-3 count kernel (a timed launch); C4: layout-3 walk kernel (a timed launch)
timeout 300 $NCU -k regex:'count3_kernel' -s 12 -o gpurun_out/${R}_count3_c2 $B --workload c2 --no-locate --no-configs > /dev/null 2> gpurun_out/${R}_prof_c.err
timeout 300 $NCU -k regex:'count3_kernel' -s 12 -o gpurun_out/${R}_count3_c5 $B --workload c5 --no-locate --no-configs > /dev/null 2> gpurun_out/${R}_prof_d.err
timeout 300 $NCU -k regex:'walk3_kernel' -s 3 -o gpurun_out/${R}_walk3_c4 $B --workload c2 --no-configs --locate-patterns 200000 > /dev/null 2> gpurun_out/${R}_prof_e.err
if [[ "${2:-}" == "all" ]]; then
  # the large-table index (count2_kernel<true,false> again: 14 default + 7 stepping launches first), and the TMA-staged variant
  timeout 300 $NCU -k regex:'count2_kernel' -s 26 -o gpurun_out/${R}_count2_c3_large $B --no-locate --no-configs > /dev/null 2> gpurun_out/${R}_prof_f.err
  CSFM_PATTERN_STAGING=tma timeout 300 $NCU -k regex:'count2_tma_kernel' -s 12 -o gpurun_out/${R}_count2_tma $B --no-locate --no-configs --no-large-table > /dev/null 2> gpurun_out/${R}_prof_g.err
fi
ls -la gpurun_out/${R}_*.ncu-rep
