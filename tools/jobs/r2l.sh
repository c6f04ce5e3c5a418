# profiles for round 2: launch list of the bench command, full captures of the step and rollout kernels, per-config table
B="python bench.py --steps 20 --warmup 5 --no-cpu --repeats 2 --burn-in 100 --clock-warm-s 0 --clock-probe-s 0 --no-graph"
$B > gpurun_out/r2l_bench_plain.json 2> gpurun_out/r2l_bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2l_launches.csv $B > gpurun_out/r2l_bench_under_ncu.json 2> gpurun_out/r2l_ncu_launch.err
python tools/run_config.py ref 6 > gpurun_out/r2l_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/r2l_prof_step python tools/run_config.py ref 6 > gpurun_out/r2l_ncu1.log 2>&1
python tools/run_config.py rollout 20 > gpurun_out/r2l_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/r2l_prof_rollout python tools/run_config.py rollout 20 > gpurun_out/r2l_ncu2.log 2>&1
python tools/run_config.py ur5 6 > gpurun_out/r2l_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/r2l_prof_ur5 python tools/run_config.py ur5 6 > gpurun_out/r2l_ncu3.log 2>&1
python tools/bench_configs.py > gpurun_out/r2l_bench_configs.txt 2>&1
cat gpurun_out/r2l_bench_configs.txt | tail -30; wc -l gpurun_out/r2l_launches.csv; ls -la gpurun_out/*.ncu-rep
