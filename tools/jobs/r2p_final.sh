# final evidence run of round 2 (one GPU)
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2p_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2p_bench_ref.json 2> gpurun_out/r2p_bench_ref.err
B="python bench.py --steps 20 --warmup 5 --no-cpu --repeats 2 --burn-in 100 --clock-warm-s 0 --clock-probe-s 0 --no-graph"
$B > gpurun_out/r2p_bench_plain.json 2> gpurun_out/r2p_bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2p_launches.csv $B > gpurun_out/r2p_bench_under_ncu.json 2> gpurun_out/r2p_ncu_launch.err
python tools/run_config.py ref 6 > gpurun_out/r2p_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/r2p_prof_step python tools/run_config.py ref 6 > gpurun_out/r2p_ncu1.log 2>&1
python tools/run_config.py rollout 20 > gpurun_out/r2p_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/r2p_prof_rollout python tools/run_config.py rollout 20 > gpurun_out/r2p_ncu2.log 2>&1
python tools/run_config.py ur5 6 > gpurun_out/r2p_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/r2p_prof_ur5 python tools/run_config.py ur5 6 > gpurun_out/r2p_ncu3.log 2>&1
python tools/bench_configs.py > gpurun_out/r2p_bench_configs.txt 2>&1
python tools/parity_soak.py --arm ref --n 16384 --steps 600 > gpurun_out/r2p_soak1.log 2>&1
python tools/parity_soak.py --arm ref --n 8192 --steps 300 --float > gpurun_out/r2p_soak2.log 2>&1
python tools/parity_soak.py --arm ur5 --n 163840 --steps 30 > gpurun_out/r2p_soak3.log 2>&1
tail -3 gpurun_out/r2p_tests.log; tail -1 gpurun_out/r2p_soak1.log gpurun_out/r2p_soak2.log gpurun_out/r2p_soak3.log; cut -c1-400 gpurun_out/r2p_bench.json
python tools/run_config.py ref 60 > gpurun_out/r2p_plain4.log 2>&1 && ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:step_kernel -s 30 -c 20 --csv --log-file gpurun_out/r2p_dram_warm.csv python tools/run_config.py ref 60 > /dev/null 2>&1
tail -3 gpurun_out/r2p_dram_warm.csv
