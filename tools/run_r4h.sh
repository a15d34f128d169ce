#!/bin/bash
# aligner: launch shape chosen by resident CTAs per SM (waves x per-utterance latency); tests, then C5 at 2000 / 1000 / 4000 utterances
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r4h_*
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_em_recipe.py -m gpu -q -x 2>&1 | tail -4 > $O/r4h_tests.txt
cat $O/r4h_tests.txt
for u in 2000 1000 4000; do
  echo "## utts=$u chosen" >> $O/r4h_align.txt
  KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --utts $u --reps 4 --check 4 2>&1 | grep -v "^khg_align_batch host" | cut -c1-420 | tail -3 >> $O/r4h_align.txt
  echo "## utts=$u NT=128 FC=32" >> $O/r4h_align.txt
  KHG_ALIGN_NT=128 KHG_ALIGN_FORCE_FC=32 KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --utts $u --reps 4 --check 4 2>&1 | grep -v "^khg_align_batch host" | cut -c1-420 | tail -3 >> $O/r4h_align.txt
done
grep -o "## utts.*\|FC [0-9]* NT [0-9]*\|search [0-9.]* ms\|\"value_device_feats\": [0-9.]*" $O/r4h_align.txt | paste -s -d' ' | sed 's/## /\n/g'
