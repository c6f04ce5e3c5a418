#!/usr/bin/env python
"""Run a few steps of one config (for ncu): python tools/run_config.py {ref|ur5|generic|rollout|rollout_ur5} [steps] [other-build.so]
(`rollout*`: three mt_rollout_random(steps) calls = three launches of the multi-step rollout kernel)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from manytor_b200 import BatchedEnvs, UR5_ARM, REFERENCE_ARM
which = sys.argv[1] if len(sys.argv) > 1 else "ref"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n = 1 << 20
lib = dict(lib_path=os.path.abspath(sys.argv[3])) if len(sys.argv) > 3 else {}
if which.startswith("rollout"):
    arm, x = (UR5_ARM, 20) if which == "rollout_ur5" else (REFERENCE_ARM, 10)
    env = BatchedEnvs(n, x, arm=arm, device=0, auto_reset=True, horizon=1000, seed=3, **lib)
    env.reset()
    for _ in range(3):
        env.rollout_random(steps)
    torch.cuda.synchronize()
    print("ok", which, env.launch_count)
    sys.exit(0)
if which == "ur5":
    env = BatchedEnvs(n, 20, arm=UR5_ARM, device=0, auto_reset=True, horizon=1000, seed=3, **lib)
elif which == "generic":
    env = BatchedEnvs(n, 10, device=0, auto_reset=True, horizon=1000, seed=3, fk_mode=1, **lib)
else:
    env = BatchedEnvs(n, 10, device=0, auto_reset=True, horizon=1000, seed=3, **lib)
env.reset()
acts = torch.randint(-180, 180, (n, env.j), device="cuda").float()
for _ in range(steps):
    env.step(acts)
torch.cuda.synchronize()
print("ok", which, env.launch_count)
