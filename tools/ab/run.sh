python tools/ab.py tools/ab/MB7.so tools/ab/MB6.so 5
python tools/ab.py tools/ab/MB7.so tools/ab/MB5.so 5
