#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2r_*.txt
timeout 600 python -m pytest tests/test_gpu_stats_tc.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > $O/r2r_tests.txt
for r in 1 2; do
for c in c4 c5; do echo "new $c" >> $O/r2r_bench_stats.txt; timeout 120 python tools/bench_stats.py $c >> $O/r2r_bench_stats.txt 2>&1; 
echo "prev $c" >> $O/r2r_bench_stats.txt; KHG_B200_LIB=tools/ab/stk_prev.so timeout 120 python tools/bench_stats.py $c >> $O/r2r_bench_stats.txt 2>&1; done; done
tail -4 $O/r2r_tests.txt; cut -c1-160 $O/r2r_bench_stats.txt
