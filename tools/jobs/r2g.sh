python tools/run_config.py ref 6 > gpurun_out/r2g_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/r2g_prof_head python tools/run_config.py ref 6 > gpurun_out/r2g_ncu1.log 2>&1
python tools/run_config.py ref 6 build/variants/r2opt1.so > gpurun_out/r2g_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/r2g_prof_opt1 python tools/run_config.py ref 6 build/variants/r2opt1.so > gpurun_out/r2g_ncu2.log 2>&1
tail -2 gpurun_out/r2g_ncu1.log gpurun_out/r2g_ncu2.log
