import sys, os, time, json
sys.path.insert(0, os.getcwd())
import torch
from manytor_b200 import BatchedEnvs
PEAK=6453.1
n=1<<20
for x in (5, 7, 8, 12, 10):
    env = BatchedEnvs(n, x, device=0, auto_reset=True, horizon=1000, seed=3)
    env.reset(); env.rollout_random(300, write_obs=False)
    acts=[torch.randint(-180,180,(n,4),device="cuda").float() for _ in range(8)]
    for _ in range(100): env.step(acts[0])
    torch.cuda.synchronize()
    ts=[]
    for r in range(4):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200): env.step(acts[i&7])
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)/200)
    ms=sorted(ts)[len(ts)//2]; B=env.bytes_per_env_step(True,True)
    print(f"x={x:2d}  {ms*1e3:7.2f} us/step  {n/ms*1e3:.3e} env-steps/s  {B} B  {n*B/ms/1e6/PEAK*100:5.1f}% of roofline")
    env.close()
