#!/bin/bash
# host -> device staging of pageable buffers by several threads: host-path tests, then W-aligned and the aligner from numpy inputs
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r4p_*
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_align.py tests/test_reference_api.py tests/test_gselect.py -m gpu -q -x 2>&1 | tail -3 > $O/r4p_tests.txt
cat $O/r4p_tests.txt
for th in 1 default; do
  echo "## KHG_STAGE_THREADS=$th" >> $O/r4p_stage.txt
  if [ $th = 1 ]; then export KHG_STAGE_THREADS=1; else unset KHG_STAGE_THREADS; fi
  KHG_BENCH_HOST=1 timeout 200 python tools/bench_stats.py c4 2>&1 | tail -2 >> $O/r4p_stage.txt
  timeout 300 python tools/bench_align.py --reps 4 --check 0 2>&1 | tail -1 | grep -o '"value_device_feats": [0-9.]*\|"e2e_host_feats": [0-9.]*' | paste -s -d' ' >> $O/r4p_stage.txt
done
cat $O/r4p_stage.txt | cut -c1-250
