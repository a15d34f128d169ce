#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r4b_*.txt
timeout 600 python -m pytest tests/test_gpu_stats_tc.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 > $O/r4b_tests.txt
for c in c3 c2; do timeout 120 python tools/stats_tc_check.py $c 400000 2>&1 | tail -2 >> $O/r4b_tests.txt; done
for r in 1 2; do
for c in c4 c5 c3 c2; do echo "new $c" >> $O/r4b_bench_stats.txt; timeout 120 python tools/bench_stats.py $c >> $O/r4b_bench_stats.txt 2>&1; 
echo "prev $c" >> $O/r4b_bench_stats.txt; KHG_B200_LIB=tools/ab/stk_prev.so timeout 120 python tools/bench_stats.py $c >> $O/r4b_bench_stats.txt 2>&1; done; done
cut -c1-300 $O/r4b_tests.txt; grep -o 'new c.\|prev c.\|"frames_per_s": [0-9.]*' $O/r4b_bench_stats.txt | paste - -
