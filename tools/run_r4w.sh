#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_reference_api.py -m gpu -q -x 2>&1 | tail -4 > $O/r4w_tests.txt
cat $O/r4w_tests.txt
timeout 300 python tools/latency_probe.py > $O/r4w_latency.txt 2>&1; cat $O/r4w_latency.txt
KHG_STAGE_THREADS=1 timeout 300 python tools/latency_probe.py > $O/r4w_latency_1thread.txt 2>&1; cat $O/r4w_latency_1thread.txt
