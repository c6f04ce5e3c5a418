python tools/ab.py tools/ab/A1_packed_nohint.so tools/ab/B_l2hint.so 6
python tools/ab.py tools/ab/A0_scalar_persistent.so tools/ab/A1_packed_nohint.so 4
