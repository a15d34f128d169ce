set -x
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json
ncu --set full --clock-control none --import-source on -k regex:loglikes_tc -s 2 -c 1 -f -o gpurun_out/prof_tc_r1h python tools/prof_dense.py > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_tc_r1h.ncu-rep > gpurun_out/r1h_ncu_loglikes_tc_f16.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1h_launches.csv python bench.py --frames 3000000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/r1h_ncu_loglikes_tc_f16.txt
