#!/usr/bin/env python
"""Long lock-step parity run of the CUDA path against the fp64 oracle (tests/parity.py machinery):
python tools/parity_soak.py [n_envs] [steps]  -> gpurun_out/parity_soak.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from manytor_b200 import BatchedEnvs
from oracle import OracleEnvs, REFERENCE_ARM, sample_points_reference_stream
from parity import lockstep

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
float_actions = len(sys.argv) > 3 and sys.argv[3] == "float"      # continuous targets instead of action_sample()'s integers
np.random.seed(2024)
pts = np.stack([sample_points_reference_stream(10) for _ in range(n)])      # the reference's RNG
env = BatchedEnvs(n, 10, device=0)
ora = OracleEnvs(n, 10)
ora.reset(points=pts); env.reset(); env.set_points(pts)
rng = np.random.RandomState(7); fresh_rng = np.random.RandomState(8)

def fresh(t, done):
    out = np.zeros((n, 10, 3))
    for i in np.nonzero(done)[0]:
        k = 0
        while k < 10:
            c = fresh_rng.uniform(-51.3, 51.3, size=3)
            if c[2] >= 0 and np.sqrt((c ** 2).sum()) <= 51.3:
                out[i, k] = c; k += 1
    return out

t0 = time.time()
act = (lambda t: rng.uniform(-180, 180, size=(n, 4))) if float_actions else (lambda t: rng.randint(-180, 180, size=(n, 4)))
rep = lockstep(env, ora, REFERENCE_ARM, act, steps, on_done=fresh)
out = dict(env_steps=rep.env_steps, max_joint_err_abs=rep.max_joint_err, max_joint_err_rel_reach=rep.max_joint_err / 55.6,
           max_dist_err=rep.max_dist_err, max_angle_err_over_allowed=rep.max_angle_excess,
           near_threshold_flips=rep.near_threshold, hard_mismatches=rep.hard_mismatch, notes=rep.notes[:5],
           seconds=time.time() - t0, config=f"{n} envs x {steps} steps, x=10, {'continuous' if float_actions else 'integer'} actions, objectives from the reference RNG, refresh on done")
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_soak.json"), "w"), indent=1)
