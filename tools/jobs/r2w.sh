L=manytor_b200/lib/libmanytor_b200.so
python tools/ab.py --isolate 2 --rounds 5 --steps 400 --modes step $L $L@MT_POL_STORE=normal $L@MT_POL_LOAD=normal $L@MT_POL_LOAD=normal,MT_POL_STORE=normal $L@MT_POL_STORE=last > gpurun_out/r2w_ab.txt 2>&1
cat gpurun_out/r2w_ab.txt
