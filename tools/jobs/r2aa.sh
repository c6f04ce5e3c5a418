python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_shape.py -m gpu -q -x -k "ur5 or custom or generic or substep or selectors or nvrtc or runtime" 2>&1 | tail -3
python tools/ab.py --isolate 2 --rounds 5 --steps 300 --modes step,rand --arm ur5 --x 20 build/variants/r2head.so manytor_b200/lib/libmanytor_b200.so > gpurun_out/r2aa_ab_ur5.txt 2>&1
python tools/ab.py --isolate 1 --rounds 5 --steps 300 --modes step --arm ur5 --x 20 --fk-mode 1 build/variants/r2head.so manytor_b200/lib/libmanytor_b200.so >> gpurun_out/r2aa_ab_ur5.txt 2>&1
cat gpurun_out/r2aa_ab_ur5.txt
