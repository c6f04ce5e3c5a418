python tools/ab.py --isolate 2 --rounds 5 --steps 400 --modes step,rand build/variants/r2opt1.so manytor_b200/lib/libmanytor_b200.so build/variants/v2_MT_V_EPRUNTIME.so build/variants/v2_MT_V_D3PACKED.so build/variants/v2_MT_V_ATAN7.so > gpurun_out/r2e_ab.txt 2>&1
cat gpurun_out/r2e_ab.txt
