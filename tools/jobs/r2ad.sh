L=manytor_b200/lib/libmanytor_b200.so
python tools/ab.py --isolate 1 --rounds 5 --steps 100 --modes step --lg 22 $L $L@MT_L2_KEEP_MB=0 $L@MT_L2_KEEP_MB=72 $L@MT_L2_KEEP_MB=0,MT_POL_LOAD=normal,MT_POL_STORE=normal $L@MT_POL_STORE=normal $L@MT_WARPS_PER_BLOCK=24 > gpurun_out/r2ad_ab.txt 2>&1
cat gpurun_out/r2ad_ab.txt
