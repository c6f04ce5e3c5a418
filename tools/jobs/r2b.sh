ls -la oracle/_ref > gpurun_out/r2b_ref_ls.txt 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_shape.py tests/test_gpu_round2.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r2b_tests.log
python tools/ab.py --isolate 2 --rounds 5 --steps 400 build/variants/r2base.so manytor_b200/lib/libmanytor_b200.so > gpurun_out/r2b_ab.txt 2>&1
python tools/run_config.py ref 6 > gpurun_out/r2b_plain.log 2>&1 && ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:step_kernel -s 2 -c 2 --csv --log-file gpurun_out/r2b_inst_new.csv python tools/run_config.py ref 6 > /dev/null 2>&1
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:step_kernel -s 2 -c 2 --csv --log-file gpurun_out/r2b_inst_base.csv python tools/run_config.py ref 6 build/variants/r2base.so > /dev/null 2>&1
tail -8 gpurun_out/r2b_tests.log; cat gpurun_out/r2b_ab.txt; tail -3 gpurun_out/r2b_inst_new.csv gpurun_out/r2b_inst_base.csv
