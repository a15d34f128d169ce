#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2x_*.txt
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -12 > $O/r2x_tests.txt
KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --out $O/r2x_align_c5.json > $O/r2x_align_c5.txt 2>&1
KHG_ALIGN_TILE_SUBSET=0 KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --reps 3 --check 2 > $O/r2x_align_c5_full.txt 2>&1
tail -8 $O/r2x_tests.txt; tail -5 $O/r2x_align_c5.txt | cut -c1-600; tail -3 $O/r2x_align_c5_full.txt | cut -c1-600
