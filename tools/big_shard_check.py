import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from manytor_b200 import BatchedEnvs
for n in (1 << 24, (1 << 25) + 12345):
    env = BatchedEnvs(n, 10, device=0, auto_reset=True, horizon=20, seed=3)
    env.reset()
    pos = neg = 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(25):
        obs, rew, done = env.rollout_random(1)
        pos += int((rew == 1).sum()); neg += int((rew == -1).sum())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    s = env.stats()
    assert s["env_steps"] == 25 * n and s["episodes"] >= n
    assert s["reward_sum"] + s["live_reward_sum"] == pos - neg, (s, pos, neg)
    assert bool(torch.isfinite(obs).all()) and bool((obs[-1] >= 0).all())
    print(n, "ok", s["episodes"], f"{25*n/dt:.3e} env-steps/s incl. host reductions", torch.cuda.max_memory_allocated() >> 20, "MiB torch")
    env.close(); del env, obs, rew, done
    torch.cuda.empty_cache()
