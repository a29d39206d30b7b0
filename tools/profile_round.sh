#!/usr/bin/env bash
# Round-end evidence on a B200 (run through gpurun): each command first runs WITHOUT ncu (must exit 0), then the
# launch list of the default bench and one `ncu --set full` capture per kernel of interest. Launch order of the
# count kernel in a `bench.py --steps 3 --warmup 3` count leg: 8 instrumented launches (<.,true>), 3 warm-up, 3 timed.
set -x
mkdir -p gpurun_out
R=${1:-r2}
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 300 $B --locate-patterns 200000 > gpurun_out/${R}_prof_plain.json 2> gpurun_out/${R}_prof_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'count[23]|walk[23]|single|widen|narrow|expand_rows|rows_to|DeviceScan|kmer|dna_|nib_' -c 900 --csv \
  --log-file gpurun_out/${R}_launches_bench_default.csv $B --locate-patterns 200000 > gpurun_out/${R}_prof_ll.json 2> gpurun_out/${R}_prof_ll.err
# C3: default kernel (count2<true,false>) and the stepping kernel (count2<false,false>)
timeout 300 $NCU -k regex:'count2_kernel' -s 12 -o gpurun_out/${R}_count2_c3_default $B --no-locate --no-configs --no-large-table > /dev/null 2> gpurun_out/${R}_prof_a.err
timeout 300 $NCU -k regex:'count2_kernelILb0ELb0' -s 3 -o gpurun_out/${R}_count2_c3_stepping $B --no-locate --no-configs --no-large-table > /dev/null 2> gpurun_out/${R}_prof_b.err
# C2 / C5: layout-3 count kernel; C4: layout-3 walk kernel
timeout 300 $NCU -k regex:'count3_kernelILb0' -s 4 -o gpurun_out/${R}_count3_c2 $B --workload c2 --no-locate --no-configs > /dev/null 2> gpurun_out/${R}_prof_c.err
timeout 300 $NCU -k regex:'count3_kernelILb0' -s 4 -o gpurun_out/${R}_count3_c5 $B --workload c5 --no-locate --no-configs > /dev/null 2> gpurun_out/${R}_prof_d.err
timeout 300 $NCU -k regex:'walk3_kernel' -s 3 -o gpurun_out/${R}_walk3_c4 $B --workload c2 --no-configs --locate-patterns 200000 > /dev/null 2> gpurun_out/${R}_prof_e.err
ls -la gpurun_out/*.ncu-rep
