#!/usr/bin/env python
"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from manytor_b200 import BatchedEnvs, UR5_ARM

for n, x, arm, kw in ((200, 10, None, {}), (97, 7, None, {}), (130, 20, UR5_ARM, {}), (64, 10, None, dict(fk_mode=1))):
    args = dict(device=0, auto_reset=True, horizon=3, seed=1, **kw)
    if arm is not None:
        args["arm"] = arm
    env = BatchedEnvs(n, x, **args)
    env.reset()
    for _ in range(6):
        a = env.sample_actions()
        env.step(a, joints=True)
        env.rollout_random(1)
        env.rollout_random(1, write_obs=False)
    env.observe(); env.get_points(); env.get_state(); env.stats(); env.fetch_env(3)
    env.set_points(np.zeros((n, x, 3), dtype=np.float32), mask=np.arange(n) % 2 == 0)
    env.step_host(np.zeros((n, env.j), dtype=np.float32))
    torch.cuda.synchronize()
    env.close()
print("sanitize smoke done")
