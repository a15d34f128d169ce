#!/bin/bash
# compute-sanitizer over the kernels added in round 2: K3t (tensor-core statistics), merge_virtual_kernel, the tile-subset dense kernel
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2_sanitizer.txt
echo "## memcheck: tests/test_gpu_stats_tc.py" >> $O/r2_sanitizer.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_stats_tc.py -q -x 2>&1 | tail -6 >> $O/r2_sanitizer.txt
echo "## memcheck: big pdfs / chunked scratch (tests/test_gpu_tc.py -k '240 or ubm or scratch')" >> $O/r2_sanitizer.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_tc.py -q -x -k "240 or ubm or scratch" 2>&1 | tail -6 >> $O/r2_sanitizer.txt
echo "## memcheck: aligner tile subset" >> $O/r2_sanitizer.txt
KHG_ALIGN_EXACT=none timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_align.py -q -x -k "tiles_the_graphs_need and none" 2>&1 | tail -6 >> $O/r2_sanitizer.txt
echo "## racecheck: K3t (one BASELINE-shaped case)" >> $O/r2_sanitizer.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_stats_tc.py -q -x -k "vs_oracle and 13-7" 2>&1 | tail -6 >> $O/r2_sanitizer.txt
echo "## synccheck: K3t" >> $O/r2_sanitizer.txt
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_stats_tc.py -q -x -k "vs_oracle and 13-7" 2>&1 | tail -6 >> $O/r2_sanitizer.txt
cat $O/r2_sanitizer.txt
