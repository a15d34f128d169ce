#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2z_*.txt
timeout 900 python -m pytest tests/test_gpu_align.py -m gpu -q -x 2>&1 | tail -8 > $O/r2z_tests.txt
KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --out $O/r2z_align_c5.json > $O/r2z_align_c5.txt 2>&1
timeout 300 python tools/bench_align.py --reps 5 --check 2 > $O/r2z_align_c5_untimed.txt 2>&1
tail -4 $O/r2z_tests.txt; grep "khg_align_batch:" $O/r2z_align_c5.txt | sed 's/.*chunks 1 | //' | head -6; tail -1 $O/r2z_align_c5.txt | cut -c1-500; tail -1 $O/r2z_align_c5_untimed.txt | cut -c150-500
