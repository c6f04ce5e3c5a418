L=manytor_b200/lib/libmanytor_b200.so
MT_TILE_POOL=8 timeout 600 python -m pytest tests/test_gpu_bench_shape.py -m gpu -x -q > gpurun_out/r2bb_tests.txt 2>&1
tail -3 gpurun_out/r2bb_tests.txt
python tools/ab.py --isolate 2 --rounds 5 --steps 400 --modes step build/variants/head.so $L $L@MT_TILE_POOL=4 $L@MT_TILE_POOL=8 $L@MT_TILE_POOL=12 > gpurun_out/r2bb_ab.txt 2>&1
cat gpurun_out/r2bb_ab.txt
