L=manytor_b200/lib/libmanytor_b200.so
MT_TILE_POOL=12 timeout 600 python -m pytest tests/test_gpu_bench_shape.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2ba_tests.txt 2>&1
tail -5 gpurun_out/r2ba_tests.txt
python tools/ab.py --isolate 2 --rounds 5 --steps 400 --modes step $L $L@MT_TILE_POOL=6 $L@MT_TILE_POOL=12 $L@MT_TILE_POOL=20 > gpurun_out/r2ba_ab.txt 2>&1
cat gpurun_out/r2ba_ab.txt
