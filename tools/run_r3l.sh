#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r3l_chunks.txt
for c in 303104 606208 1212416 303104 1212416; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-workloads --no-cpu --chunk $c 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk', d['config']['dense_block_frames'], 'value', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), 'k1', round(d['roofline']['dense_frames_per_s']/1e6,2), 'share', round(d['roofline']['share_of_step'],4), 'mhz', d['clocks']['sm_mhz'], d['parity_check']['ok'])
" >> $O/r3l_chunks.txt
done
cat $O/r3l_chunks.txt
