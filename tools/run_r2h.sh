#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "not align" 2>&1 | tail -15 > $O/r2h_tests.txt
for c in c4 c3 c2 c5; do timeout 120 python tools/bench_stats.py $c >> $O/r2h_bench_stats.txt 2>&1; done
for k in 1 2; do echo "ctas_per_sm $k" >> $O/r2h_bench_stats.txt; KHG_STATS_CTAS_PER_SM=$k timeout 120 python tools/bench_stats.py c4 >> $O/r2h_bench_stats.txt 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stats_kernel -s 1 -c 1 -o $O/r2h_stats python tools/prof_dense.py 2000000 1 > $O/r2h_ncu.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "acc_stats or stats" 2>&1 | tail -8 > $O/r2h_sanitizer.txt
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "acc_stats_ali_vs_oracle" 2>&1 | tail -8 >> $O/r2h_sanitizer.txt
tail -6 $O/r2h_tests.txt; cat $O/r2h_bench_stats.txt | cut -c1-200; tail -12 $O/r2h_sanitizer.txt
