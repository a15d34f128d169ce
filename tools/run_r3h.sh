#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_stats_tc.py -m gpu -q 2>&1 | tail -25 > $O/r3h_tests.txt
tail -25 $O/r3h_tests.txt
