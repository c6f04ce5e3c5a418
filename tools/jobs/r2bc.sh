L=manytor_b200/lib/libmanytor_b200.so
MT_TILE_POOL=12 timeout 600 python -m pytest tests/test_gpu_bench_shape.py -m gpu -x -q -k "reference_arm" > gpurun_out/r2bc_tests.txt 2>&1
tail -3 gpurun_out/r2bc_tests.txt
python tools/ab.py --isolate 2 --rounds 5 --steps 400 --modes step build/variants/head.so build/variants/nopool.so $L $L@MT_TILE_POOL=6 $L@MT_TILE_POOL=12 $L@MT_TILE_POOL=25 > gpurun_out/r2bc_ab.txt 2>&1
cat gpurun_out/r2bc_ab.txt
