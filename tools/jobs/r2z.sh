L=manytor_b200/lib/libmanytor_b200.so
python tools/ab.py --isolate 2 --rounds 5 --steps 400 --modes step,rand $L $L@MT_WARPS_PER_BLOCK=14 $L@MT_WARPS_PER_BLOCK=24 $L@MT_WARPS_PER_BLOCK=20 > gpurun_out/r2z_ab.txt 2>&1
cat gpurun_out/r2z_ab.txt
