#!/usr/bin/env bash
# Round-end evidence in one gpurun call: the GPU suite, the bounds-check build over the parity and the full-size suites,
# the ncu round (tools/profile_round.sh) and the default bench line. Afterwards, here: tools/profile_collect.sh <R>.
R=${1:-r2k}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/${R}_pytest_gpu_full.log 2>&1; tail -1 gpurun_out/${R}_pytest_gpu_full.log
python tools/bounds_check.py > gpurun_out/${R}_bounds_parity.out 2>&1; tail -2 gpurun_out/${R}_bounds_parity.out
cp gpurun_out/bounds_check.log gpurun_out/${R}_bounds_check_parity.log
python tools/bounds_check.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_c3_c4.py tests/test_gpu_fullsize_c5.py tests/test_gpu_sa_builder.py -x -q -m gpu > gpurun_out/${R}_bounds_full.out 2>&1; tail -2 gpurun_out/${R}_bounds_full.out
cp gpurun_out/bounds_check.log gpurun_out/${R}_bounds_check_full.log
bash tools/profile_round.sh $R all > gpurun_out/${R}_profile_round.out 2>&1; tail -9 gpurun_out/${R}_profile_round.out
python bench.py > gpurun_out/${R}_bench_default_n1.json 2> gpurun_out/${R}_bench_default_n1.err; python -c "
import json; d = json.load(open('gpurun_out/${R}_bench_default_n1.json')); print('%.3e q/s e2e %.3e' % (d['value'], d['e2e']['value']), 'locate %.3e' % d['locate']['value'], 'c2 %.3e' % d['configs']['c2']['value'], 'c5 %.3e' % d['configs']['c5']['value'])"
