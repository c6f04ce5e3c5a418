python tools/ab.py tools/ab/base.so tools/ab/pts_normal.so 5 | grep "step("
python tools/ab.py tools/ab/base.so tools/ab/obs_normal.so 5 | grep "step("
python tools/ab.py tools/ab/base.so tools/ab/act_keep.so 5 | grep "step("
