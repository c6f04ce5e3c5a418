L=manytor_b200/lib/libmanytor_b200.so
python tools/ab.py --isolate 2 --rounds 5 --steps 400 --modes step $L $L@MT_L2_KEEP_MB=0 $L@MT_L2_KEEP_MB=64 $L@MT_L2_KEEP_MB=96 $L@MT_POL_STORE=first > gpurun_out/r2y_ab.txt 2>&1
python tools/ab.py --isolate 1 --rounds 5 --steps 200 --modes step --lg 21 $L $L@MT_POL_LOAD=normal,MT_POL_STORE=normal $L@MT_POL_LOAD=normal,MT_POL_STORE=normal,MT_L2_KEEP_MB=48 $L@MT_L2_KEEP_MB=48 >> gpurun_out/r2y_ab.txt 2>&1
cat gpurun_out/r2y_ab.txt
