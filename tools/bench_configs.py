#!/usr/bin/env python
"""Throughput of every BASELINE.json config that runs on one GPU (configs 2, 3, 5), all modes.
Device time with CUDA events, 1 s clock warm-up, median of 5 rounds of 300 steps (mt_step: a 300-node CUDA graph)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from manytor_b200 import ArmSpec, BatchedEnvs, UR5_ARM, REFERENCE_ARM

PEAK = 6453.1
# UR5 with other link lengths: not a built-in preset, so it goes through the NVRTC path
CUSTOM6 = ArmSpec(dh=tuple((r[0] * 1.1, r[1], r[2] * 0.9, r[3]) for r in UR5_ARM.dh), obs_frame=5, ground_frames=(5, 6),
                  catch_frame=6, radius=0.85, catch_tol=0.13)
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def bench(name, n, x, arm, K=300, rounds=5, **kw):
    env = BatchedEnvs(n, x, arm=arm, device=0, auto_reset=True, horizon=1000, seed=3, **kw)
    env.reset()
    env.rollout_random(1000, write_obs=False)
    J = arm.n_joints
    acts = [torch.randint(-180, 180, (n, J), device="cuda").float() for _ in range(8)]
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 1.0:
        env.rollout_random(100)
        torch.cuda.synchronize()
    # K steps with their own action buffers as ONE CUDA graph, like bench.py: Python cannot issue a launch
    # every ~45 us, and a per-call timing would measure the interpreter
    env.step(acts[0])
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(K):
            env.step(acts[i & 7])
    graph.replay()
    torch.cuda.synchronize()
    out = {}
    for mode, fn, hbm_act, wobs in (("step(actions in HBM)", graph.replay, True, True),
                                    ("rollout_random", lambda: env.rollout_random(K), False, True),
                                    ("rollout_random, no obs", lambda: env.rollout_random(K, write_obs=False), False, False)):
        ts = []
        for _ in range(rounds):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / K)
        ms = sorted(ts)[len(ts) // 2]
        B = env.bytes_per_env_step(hbm_act, wobs)
        out[mode] = dict(us_per_step=ms * 1e3, env_steps_per_s=n / ms * 1e3, bytes_per_env_step=B,
                         gbs=n * B / ms / 1e6, frac=n * B / ms / 1e6 / PEAK)
        print(f"{name:44s} {mode:24s} {ms*1e3:8.2f} us  {n/ms*1e3:.3e} env-steps/s  {B} B  "
              f"{n*B/ms/1e6:7.0f} GB/s  {n*B/ms/1e6/PEAK*100:5.1f}% of {PEAK:.0f}")
    return out


res = {}
CASES = [("config2: 4096 envs, J=4, x=10 (in L2)", 4096, 10, REFERENCE_ARM, {}),
         ("config3: 2^20 envs, J=4, x=10", 1 << 20, 10, REFERENCE_ARM, {}),
         ("config3 via generic DH chain (fk_mode=1)", 1 << 20, 10, REFERENCE_ARM, dict(fk_mode=1)),
         ("config5: 2^20 envs, J=6 UR5, x=20", 1 << 20, 20, UR5_ARM, {}),
         ("config5 run-time table (fk_mode=1)", 1 << 20, 20, UR5_ARM, dict(fk_mode=1)),
         ("6-DOF custom table, x=20, NVRTC-specialised", 1 << 20, 20, CUSTOM6, dict(fk_mode=3)),
         ("2^22 envs, J=4, x=10", 1 << 22, 10, REFERENCE_ARM, dict(K=100))]
only = sys.argv[1] if len(sys.argv) > 1 else ""
for name, n, x, arm, kw in CASES:
    if only in name:
        res[name] = bench(name, n, x, arm, **kw)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_configs.json", "w"), indent=1)
