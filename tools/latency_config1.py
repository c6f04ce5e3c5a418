#!/usr/bin/env python
"""Where the microseconds of BASELINE configs[0] go: ONE env stepped through the drop-in.
  python tools/latency_config1.py [steps]
Prints us per iteration of (a) the reference-shaped loop `env.step(env.action_sample())`, (b) `env.step` with a
fixed action, (c) `BatchedEnvs.step_host` on its own page-locked buffer, (d) the bare `mt_step_host` call through
ctypes with prebuilt arguments, (e) `action_sample()` alone."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import manytor_b200.manytor as tor

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
env = tor.Environment(10, seed=1)
env.reset()
for _ in range(50):
    env.step(env.action_sample())


def timed(fn):
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return 1e6 * (time.perf_counter() - t0) / steps


def loop():
    _, _, d = env.step(env.action_sample())
    if d:
        env.reset()


a = env.action_sample()
be = env._envs
buf = be.pinned("actions", (1, 4), np.float32)
buf[:] = a
h, lib = be._h, be._lib
io = be._host_io[True]
args = (h, io[1], io[3], io[5], io[7])
print(f"host step mode: {be.host_step_mode}")
print(f"(a) env.step(env.action_sample())   {timed(loop):7.2f} us")
print(f"(b) env.step(fixed action)          {timed(lambda: env.step(a)):7.2f} us")
print(f"(c) BatchedEnvs.step_host(pinned)   {timed(lambda: be.step_host(buf)):7.2f} us")
print(f"(d) mt_step_host via ctypes          {timed(lambda: lib.mt_step_host(*args)):7.2f} us")
print(f"(e) env.action_sample()              {timed(env.action_sample):7.2f} us")
