#!/usr/bin/env python
"""A/B timing of two builds of libmanytor_b200.so in ONE process on ONE GPU, interleaved
round by round so that clocks, temperature and the box are the same for both.
  python tools/ab.py tools/ab/A.so tools/ab/B.so [rounds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from manytor_b200 import BatchedEnvs

paths = [os.path.abspath(p) for p in sys.argv[1:3]]
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 6
n, K = 1 << 20, 400
envs = []
for p in paths:
    e = BatchedEnvs(n, 10, device=0, auto_reset=True, horizon=1000, seed=1, lib_path=p)
    e.reset()
    e.rollout_random(1000, write_obs=False)
    envs.append(e)
acts = [torch.randint(-180, 180, (n, 4), device="cuda").float() for _ in range(16)]
torch.cuda.synchronize()
t0 = time.perf_counter()
while time.perf_counter() - t0 < 1.5:
    for e in envs:
        e.rollout_random(100)
    torch.cuda.synchronize()

def run(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / K

modes = {
    "step(actions from HBM), obs": lambda e: [e.step(acts[i & 15]) for i in range(K)],
    "rollout_random, obs": lambda e: e.rollout_random(K),
    "rollout_random, no obs": lambda e: e.rollout_random(K, write_obs=False),
}
res = {m: [[] for _ in envs] for m in modes}
for r in range(rounds):
    for m, fn in modes.items():
        for i, e in enumerate(envs):
            res[m][i].append(run(lambda: fn(e)))
for m in modes:
    for i, p in enumerate(paths):
        v = sorted(res[m][i])
        print(f"{m:32s} {os.path.basename(p):12s} median {v[len(v)//2]:.2f} us/step  min {v[0]:.2f}  max {v[-1]:.2f}")
