#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r5a_*
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_em_recipe.py -m gpu -q -x 2>&1 | tail -3 > $O/r5a_tests.txt
cat $O/r5a_tests.txt
KHG_ALIGN_PREP_CACHE=0 KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --reps 4 --check 4 > $O/r5a_cold.txt 2>&1
grep -o "host prep.*\|khg_align_batch host.*" $O/r5a_cold.txt | sed -n 5,10p; tail -1 $O/r5a_cold.txt | grep -o '"value_device_feats": [0-9.]*\|"e2e_host_feats": [0-9.]*' | paste -s -d' '
for r in 1 2; do timeout 300 python tools/bench_align.py --reps 4 --check 0 2>&1 | tail -1 | grep -o '"value_device_feats": [0-9.]*\|"e2e_host_feats": [0-9.]*' | paste -s -d' '; done
KHG_ALIGN_PREP_CACHE=0 timeout 300 python tools/bench_align.py --reps 4 --check 0 2>&1 | tail -1 | grep -o '"value_device_feats": [0-9.]*\|"e2e_host_feats": [0-9.]*' | paste -s -d' '
