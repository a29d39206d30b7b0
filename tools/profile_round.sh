set -x
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 200 $B > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'count2|walk2|widen|narrow|expand_rows|rows_to|DeviceScan|kmer_build' -c 600 --csv --log-file gpurun_out/r1_v4_launches_bench_c3.csv $B > gpurun_out/prof_ll.json 2> gpurun_out/prof_ll.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:count2_kernel -s 12 -c 1 -f -o gpurun_out/count2_default $B --no-locate > /dev/null 2> gpurun_out/prof_a.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:count2_kernel -s 26 -c 1 -f -o gpurun_out/count2_large $B --no-locate > /dev/null 2> gpurun_out/prof_b.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk2_kernel -s 3 -c 1 -f -o gpurun_out/walk2 $B --no-large-table > /dev/null 2> gpurun_out/prof_c.err
ls -la gpurun_out/*.ncu-rep
