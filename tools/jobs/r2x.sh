L=manytor_b200/lib/libmanytor_b200.so
python tools/ab.py --isolate 2 --rounds 5 --steps 400 --modes step,rand $L $L@MT_POL_LOAD=normal $L@MT_POL_LOAD=normal,MT_POL_STORE=normal $L@MT_POL_LOAD=normal,MT_L2_KEEP_MB=48 $L@MT_POL_LOAD=normal,MT_L2_KEEP_MB=24 $L@MT_POL_LOAD=normal,MT_POL_STORE=normal,MT_L2_KEEP_MB=48 > gpurun_out/r2x_ab.txt 2>&1
python tools/ab.py --isolate 1 --rounds 5 --steps 300 --modes step --arm ur5 --x 20 $L $L@MT_POL_LOAD=normal $L@MT_POL_LOAD=normal,MT_POL_STORE=normal > gpurun_out/r2x_ab_ur5.txt 2>&1
python tools/ab.py --isolate 1 --rounds 5 --steps 100 --modes step --lg 22 $L $L@MT_POL_LOAD=normal $L@MT_POL_LOAD=normal,MT_POL_STORE=normal > gpurun_out/r2x_ab_2p22.txt 2>&1
cat gpurun_out/r2x_ab.txt gpurun_out/r2x_ab_ur5.txt gpurun_out/r2x_ab_2p22.txt
