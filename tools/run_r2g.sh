#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "not align" 2>&1 | tail -15 > $O/r2g_tests.txt
timeout 120 python tools/bench_stats.py > $O/r2g_bench_stats.txt 2>&1
for c in c3 c2 c5; do timeout 120 python tools/bench_stats.py $c >> $O/r2g_bench_stats.txt 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stats_kernel -s 1 -c 1 -o $O/r2g_stats python tools/prof_dense.py 2000000 1 > $O/r2g_ncu.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "acc_stats or stats" 2>&1 | tail -8 > $O/r2g_sanitizer.txt
tail -6 $O/r2g_tests.txt; cat $O/r2g_bench_stats.txt | cut -c1-260; tail -4 $O/r2g_sanitizer.txt
