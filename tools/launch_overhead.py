#!/usr/bin/env python
"""Where do the microseconds between ncu's per-launch time and the back-to-back
step rate go?  Compares one C-level loop of K launches against K Python-level calls."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from manytor_b200 import BatchedEnvs

n, K = 1 << 20, 1000
env = BatchedEnvs(n, 10, device=0, auto_reset=True, horizon=1000, seed=1)
env.reset()
env.rollout_random(1000, write_obs=False)
acts = [torch.randint(-180, 180, (n, 4), device="cuda").float() for _ in range(16)]
torch.cuda.synchronize()
t0 = time.perf_counter()
while time.perf_counter() - t0 < 1.0:
    env.rollout_random(200)
    torch.cuda.synchronize()

def timed(fn, reps=3):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = time.perf_counter()
        e0.record(); fn(); e1.record()
        t_issue = time.perf_counter() - t
        torch.cuda.synchronize()
        out.append((e0.elapsed_time(e1) * 1e3 / K, t_issue * 1e6 / K))
    return out

print("C loop   rollout_random(K)        us/step (gpu, host-issue):", timed(lambda: env.rollout_random(K)))
print("py loop  K x rollout_random(1)    us/step (gpu, host-issue):", timed(lambda: [env.rollout_random(1) for _ in range(K)]))
print("py loop  K x step(actions)        us/step (gpu, host-issue):", timed(lambda: [env.step(acts[i & 15]) for i in range(K)]))
print("C loop   rollout_random(K) no obs us/step (gpu, host-issue):", timed(lambda: env.rollout_random(K, write_obs=False)))
