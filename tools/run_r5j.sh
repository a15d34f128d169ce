#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r5j_*
timeout 900 python -m pytest tests/test_gpu_stats_tc.py tests/test_gpu_parity.py tests/test_gpu_em_recipe.py tests/test_mixup.py -m gpu -q -x 2>&1 | tail -4 > $O/r5j_tests.txt
cat $O/r5j_tests.txt
for c in c4 c5; do echo "new $c" >> $O/r5j_bench_stats.txt; timeout 120 python tools/bench_stats.py $c >> $O/r5j_bench_stats.txt 2>&1; echo "prev $c" >> $O/r5j_bench_stats.txt; KHG_B200_LIB=tools/ab/stk_prev.so timeout 120 python tools/bench_stats.py $c >> $O/r5j_bench_stats.txt 2>&1; done
grep -o 'new c.\|prev c.\|"frames_per_s": [0-9.]*' $O/r5j_bench_stats.txt | paste - -
