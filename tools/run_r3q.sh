#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r3q_*.txt
timeout 600 python -m pytest tests/test_gpu_stats_tc.py -m gpu -q -x 2>&1 | tail -3 > $O/r3q_tests.txt
for r in 1 2; do
for c in c4 c5; do
echo "bulk-prefetch $c" >> $O/r3q_bench_stats.txt; timeout 120 python tools/bench_stats.py $c >> $O/r3q_bench_stats.txt 2>&1
echo "line-prefetch $c" >> $O/r3q_bench_stats.txt; KHG_B200_LIB=tools/ab/pf1.so timeout 120 python tools/bench_stats.py $c >> $O/r3q_bench_stats.txt 2>&1
echo "no-prefetch $c" >> $O/r3q_bench_stats.txt; KHG_B200_LIB=tools/ab/pf0.so timeout 120 python tools/bench_stats.py $c >> $O/r3q_bench_stats.txt 2>&1
done; done
cat $O/r3q_tests.txt; grep -o '^[a-z-]* c.\|"frames_per_s": [0-9.]*' $O/r3q_bench_stats.txt | paste - -
