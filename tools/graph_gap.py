#!/usr/bin/env python
"""Stream launches vs one CUDA graph for K steps of mt_step, alternating rounds in one process."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from manytor_b200 import BatchedEnvs

n = 1 << 20
K = int(sys.argv[1]) if len(sys.argv) > 1 else 500
env = BatchedEnvs(n, 10, device=0, auto_reset=True, horizon=1000, seed=1)
env.reset()
env.rollout_random(1000, write_obs=False)
acts = [torch.randint(-180, 180, (n, 4), device="cuda").float() for _ in range(16)]
torch.cuda.synchronize()
t0 = time.perf_counter()
while time.perf_counter() - t0 < 1.5:
    env.rollout_random(100); torch.cuda.synchronize()

def once(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / K

s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(3):
        env.step(acts[i])
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(K):
        env.step(acts[i & 15])
g.replay(); torch.cuda.synchronize()
d, gr = [], []
for r in range(6):
    d.append(once(lambda: [env.step(acts[i & 15]) for i in range(K)]))
    gr.append(once(lambda: g.replay()))
print(f"K={K} stream launches us/step:", " ".join(f"{v:.2f}" for v in d))
print(f"K={K} graph replay    us/step:", " ".join(f"{v:.2f}" for v in gr))
