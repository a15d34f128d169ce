#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2v_*.txt
timeout 600 python -m pytest tests/test_gpu_stats_tc.py -m gpu -q -x 2>&1 | tail -3 > $O/r2v_tests.txt
KHG_STATS_TC_CTAS_PER_SM=1 KHG_B200_LIB=tools/ab/stk_timing.so timeout 120 python tools/bench_stats.py c4 2>&1 | grep -v "^{" | head -2 >> $O/r2v_timing.txt
KHG_B200_LIB=tools/ab/stk_timing.so timeout 120 python tools/bench_stats.py c4 2>&1 | grep -v "^{" | head -2 >> $O/r2v_timing.txt
for r in 1 2; do
for c in c4 c5; do echo "new $c" >> $O/r2v_bench_stats.txt; timeout 120 python tools/bench_stats.py $c >> $O/r2v_bench_stats.txt 2>&1; 
echo "prev $c" >> $O/r2v_bench_stats.txt; KHG_B200_LIB=tools/ab/stk_prev.so timeout 120 python tools/bench_stats.py $c >> $O/r2v_bench_stats.txt 2>&1; done; done
cat $O/r2v_tests.txt $O/r2v_timing.txt; cut -c1-150 $O/r2v_bench_stats.txt
