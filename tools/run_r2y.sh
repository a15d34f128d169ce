#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_align.py -m gpu -q -x -k "tiles_the_graphs_need or several_chunks or feeds_acc" 2>&1 | tail -30 > $O/r2y_tests.txt
tail -6 $O/r2y_tests.txt
