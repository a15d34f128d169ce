#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/r2u_gputests.txt
tail -12 $O/r2u_gputests.txt
