#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r3p_*.txt
for c in c4 c5 c3 c2; do echo "tc $c" >> $O/r3p_bench_stats.txt; timeout 120 python tools/bench_stats.py $c >> $O/r3p_bench_stats.txt 2>&1; done
echo "fp32 c4" >> $O/r3p_bench_stats.txt; KHG_STATS_KERNEL=simt timeout 120 python tools/bench_stats.py c4 >> $O/r3p_bench_stats.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stats_tc_kernel -s 2 -c 1 -o $O/r3p_stats_tc python tools/bench_stats.py c4 2000000 > $O/r3p_ncu.log 2>&1
grep -o 'tc c.\|fp32 c.\|"frames_per_s": [0-9.]*' $O/r3p_bench_stats.txt | paste - -
