python tools/run_config.py ref 6 > gpurun_out/r2c_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/r2c_prof python tools/run_config.py ref 6 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
