#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_em_recipe.py tests/test_reference_api.py tests/test_gpu_mstep.py tests/test_mixup.py tests/test_gpu_parity.py -q 2>&1 | tail -30 > $O/r2f_tests.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stats_kernel -s 1 -c 1 -o $O/r2f_stats python tools/prof_dense.py 2000000 1 > $O/r2f_ncu.log 2>&1
timeout 120 python tools/bench_stats.py > $O/r2f_bench_stats.txt 2>&1
tail -8 $O/r2f_tests.txt; tail -2 $O/r2f_ncu.log; cat $O/r2f_bench_stats.txt
