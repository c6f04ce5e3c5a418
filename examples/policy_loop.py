#!/usr/bin/env python
"""The step after the hot path: a policy network on the same GPU consuming the
observations the step kernel wrote, with no host round trip and no copy.

    python examples/policy_loop.py [n_envs] [steps]

obs (N, 3X) fp32 is a torch view of the kernel's output buffer; actions (N, 4) fp32 is written by the
policy straight into the buffer the next step reads.  The whole iteration (policy forward + env step) is
also captured into one CUDA graph, which is how a launch-bound RL inner loop should be driven on B200.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from manytor_b200 import ManyTorVectorEnv


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    env = ManyTorVectorEnv(n, 10, max_episode_steps=200, seed=1)
    torch.manual_seed(0)
    policy = torch.nn.Sequential(torch.nn.Linear(30, 64), torch.nn.Tanh(), torch.nn.Linear(64, 4), torch.nn.Tanh()).cuda()
    obs, _ = env.reset(seed=1)
    actions = torch.empty((n, 4), device="cuda")
    ret = torch.zeros(n, device="cuda")

    def iteration():
        with torch.no_grad():
            torch.mul(policy(obs * (1.0 / 90.0)), 180.0, out=actions)     # joint targets in degrees
        o, r, term, trunc, _ = env.step(actions)                          # o IS `obs`: reset() and step() return the same storage
        ret.add_(r)
        return o

    first = obs.clone()
    for _ in range(5):
        o = iteration()
    assert o.data_ptr() == obs.data_ptr(), "reset() and step() must hand out the same observation buffer"
    assert not torch.equal(first, obs), "the policy input did not change between iterations"
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        iteration()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"eager : {n} envs x {steps} steps with an MLP policy: {n * steps / dt:.3e} env-steps/s ({dt / steps * 1e6:.1f} us/iteration)")

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            iteration()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        iteration()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        g.replay()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"graph : {n} envs x {steps} steps with an MLP policy: {n * steps / dt:.3e} env-steps/s ({dt / steps * 1e6:.1f} us/iteration)")
    st = env.episode_statistics()
    assert torch.isfinite(ret).all() and st["env_steps"] > 0
    print("episode statistics:", st, " mean return per env so far:", float(ret.mean()))


if __name__ == "__main__":
    main()
