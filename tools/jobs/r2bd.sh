python tools/trace_warps.py build/variants/trace.so 20 gpurun_out/trace_a.npy > gpurun_out/trace_a.txt 2>&1
python tools/trace_warps.py build/variants/trace.so 20 gpurun_out/trace_b.npy > gpurun_out/trace_b.txt 2>&1
tail -4 gpurun_out/trace_a.txt gpurun_out/trace_b.txt
