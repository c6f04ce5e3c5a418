#!/usr/bin/env python
"""Long lock-step parity run of the CUDA path against the fp64 oracle (tests/parity.py machinery):

    python tools/parity_soak.py [--arm ref|ur5] [--n N] [--steps S] [--float] [--out NAME]
    -> gpurun_out/parity_soak_NAME.json

ref: objectives from the reference's RNG, refreshed on done (BASELINE config 2 shape, any N);
ur5: the 6-DOF preset of BASELINE config 5 (x = 20), objectives from the on-device sampler (downloaded).
N >= 16 x 148 x 32 x 2 = 151 552 puts the 16-warp generic kernels on multi-tile warps (the benchmarked shape)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import manytor_b200
from manytor_b200 import BatchedEnvs
from oracle import REFERENCE_ARM, UR5_ARM, sample_points_reference_stream
from parity import ChunkedOracle, lockstep, reach_of

ap = argparse.ArgumentParser()
ap.add_argument("--arm", default="ref", choices=["ref", "ur5"])
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--steps", type=int, default=2000)
ap.add_argument("--float", action="store_true", help="continuous targets instead of action_sample()'s integers")
ap.add_argument("--out", default=None)
args = ap.parse_args()
n, steps = args.n, args.steps
name = args.out or f"{args.arm}_{'continuous' if args.float else 'integer'}_actions"

if args.arm == "ref":
    spec, arm, x = REFERENCE_ARM, manytor_b200.REFERENCE_ARM, 10
    np.random.seed(2024)
    pts = np.stack([sample_points_reference_stream(10) for _ in range(n)])      # the reference's RNG
    env = BatchedEnvs(n, x, device=0)
    env.reset()
    env.set_points(pts)
else:
    spec, arm, x = UR5_ARM, manytor_b200.UR5_ARM, 20
    env = BatchedEnvs(n, x, arm=arm, device=0, seed=5)
    env.reset()
    pts = env.get_points(zero_dead=False).cpu().numpy().astype(np.float64)
ora = ChunkedOracle(n, x, spec)
ora.reset(points=pts)
J = spec.n_joints
rng, fresh_rng = np.random.RandomState(7), np.random.RandomState(8)


def fresh(t, done):
    out = np.zeros((n, x, 3))
    for i in np.nonzero(done)[0]:
        k = 0
        while k < x:
            c = fresh_rng.uniform(-spec.radius, spec.radius, size=3)
            if c[2] >= 0 and np.sqrt((c ** 2).sum()) <= spec.radius:
                out[i, k] = c
                k += 1
    return out


t0 = time.time()
act = (lambda t: rng.uniform(-180, 180, size=(n, J))) if args.float else (lambda t: rng.randint(-180, 180, size=(n, J)))
rep = lockstep(env, ora, spec, act, steps, on_done=fresh)
reach = reach_of(spec)
out = dict(env_steps=rep.env_steps, max_joint_err_abs=rep.max_joint_err, max_joint_err_rel_reach=rep.max_joint_err / reach,
           max_dist_err=rep.max_dist_err, max_angle_err_over_allowed=rep.max_angle_excess,
           near_threshold_flips=rep.near_threshold, hard_mismatches=rep.hard_mismatch, notes=rep.notes[:5],
           seconds=time.time() - t0,
           config=f"{args.arm} arm, {n} envs x {steps} steps, x={x}, {'continuous' if args.float else 'integer'} actions, "
                  f"{'objectives from the reference RNG' if args.arm == 'ref' else 'objectives from the on-device sampler'}, refresh on done")
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"parity_soak_{name}.json"), "w"), indent=1)
