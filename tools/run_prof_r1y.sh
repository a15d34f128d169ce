set -x
ncu --set full --clock-control none --import-source on -k regex:loglikes_tc -s 2 -c 1 -f -o gpurun_out/prof_tc_r1y python tools/prof_dense.py > gpurun_out/ncu_full_r1y.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_tc_r1y.ncu-rep > gpurun_out/r1y_ncu_loglikes_tc_f16.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1y_launches.csv python bench.py --frames 3000000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches_r1y.log 2>&1
timeout 300 python bench.py --impl reference > gpurun_out/r1y_bench_reference_arm.json 2> gpurun_out/r1y_bench_reference_arm.err
tail -3 gpurun_out/r1y_ncu_loglikes_tc_f16.txt
cat gpurun_out/r1y_bench_reference_arm.json
