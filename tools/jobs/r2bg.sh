python tools/latency_config1.py > gpurun_out/r2bg_latency.txt 2>&1; cat gpurun_out/r2bg_latency.txt
MT_HOST_ZEROCOPY=0 python tools/latency_config1.py 2>&1 | sed 's/^/staged: /' | tee -a gpurun_out/r2bg_latency.txt
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_dropin.py tests/test_gpu_render.py -m gpu -x -q > gpurun_out/r2bg_tests.txt 2>&1
tail -3 gpurun_out/r2bg_tests.txt
