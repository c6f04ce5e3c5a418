python tools/ab.py --isolate 2 --rounds 5 --steps 400 --modes step build/variants/r2opt1.so manytor_b200/lib/libmanytor_b200.so build/variants/v3_MT_V_PREFETCH_WORDS.so build/variants/v3_all3.so > gpurun_out/r2h_ab.txt 2>&1
cat gpurun_out/r2h_ab.txt
