L=manytor_b200/lib/libmanytor_b200.so
MT_TAIL_RANKS=24 timeout 600 python -m pytest tests/test_gpu_bench_shape.py -m gpu -x -q -k "reference_arm" > gpurun_out/r2be_tests.txt 2>&1
tail -3 gpurun_out/r2be_tests.txt
python tools/ab.py --isolate 2 --rounds 5 --steps 400 --modes step build/variants/head.so $L $L@MT_TAIL_RANKS=16 $L@MT_TAIL_RANKS=24 $L@MT_TAIL_RANKS=32,MT_TAIL_RATE=512 $L@MT_TAIL_RANKS=12,MT_TAIL_RATE=128 $L@MT_TAIL_RANKS=40,MT_TAIL_RATE=512 > gpurun_out/r2be_ab.txt 2>&1
cat gpurun_out/r2be_ab.txt
