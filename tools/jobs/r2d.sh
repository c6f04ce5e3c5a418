python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_shape.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2d_tests.log
python tools/ab.py --isolate 2 --rounds 5 --steps 400 build/variants/r2base.so build/variants/r2opt1.so manytor_b200/lib/libmanytor_b200.so > gpurun_out/r2d_ab.txt 2>&1
python tools/ab.py --isolate 1 --rounds 5 --steps 300 --arm ur5 --x 20 build/variants/r2base.so manytor_b200/lib/libmanytor_b200.so > gpurun_out/r2d_ab_ur5.txt 2>&1
tail -4 gpurun_out/r2d_tests.log; cat gpurun_out/r2d_ab.txt gpurun_out/r2d_ab_ur5.txt
