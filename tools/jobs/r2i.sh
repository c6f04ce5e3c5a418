python tools/ab.py --isolate 2 --rounds 5 --steps 400 --modes step,rand build/variants/r2opt1.so build/variants/v4_pw_u1.so build/variants/v4_pw_u2.so build/variants/v4_pw_u4.so > gpurun_out/r2i_ab.txt 2>&1
cat gpurun_out/r2i_ab.txt
