python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_bench_shape.py -m gpu -q -x 2>&1 | tail -4
python tools/ab.py --isolate 2 --rounds 5 --steps 400 --modes step,rand,noobs build/variants/r2head.so manytor_b200/lib/libmanytor_b200.so > gpurun_out/r2ab_ab.txt 2>&1
python tools/ab.py --isolate 1 --rounds 5 --steps 300 --modes rand --arm ur5 --x 20 build/variants/r2head.so manytor_b200/lib/libmanytor_b200.so >> gpurun_out/r2ab_ab.txt 2>&1
cat gpurun_out/r2ab_ab.txt
