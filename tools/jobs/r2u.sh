python tools/ab.py --isolate 2 --rounds 5 --steps 400 --modes step manytor_b200/lib/libmanytor_b200.so build/variants/v5_contig.so > gpurun_out/r2u_ab.txt 2>&1
cat gpurun_out/r2u_ab.txt
