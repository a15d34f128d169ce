#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $O/r2d_gputests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2d_smoke.txt 2>&1
KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --out $O/r2d_align_c5.json > $O/r2d_align_c5.txt 2>&1
KHG_ALIGN_EXACT=all KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --reps 2 --check 2 > $O/r2d_align_c5_all_exact.txt 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 > $O/r2d_bench_n1.json 2> $O/r2d_bench_n1.err
tail -12 $O/r2d_gputests.txt; cat $O/r2d_smoke.txt | tail -3; tail -4 $O/r2d_align_c5.txt; tail -3 $O/r2d_align_c5_all_exact.txt; tail -c 400 $O/r2d_bench_n1.err
