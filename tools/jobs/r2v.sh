python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
python tools/ab.py --isolate 3 --rounds 5 --steps 400 --modes step build/variants/v6_noprefetch.so manytor_b200/lib/libmanytor_b200.so > gpurun_out/r2v_ab.txt 2>&1
cat gpurun_out/r2v_ab.txt
