#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference under fixed seeds.

Run in the build container only (``/root/reference`` is not present on the GPU
box):  ``python tests/golden/make_golden.py``.  Writes ``tests/golden/*.npz``
and prints how far the oracle (``oracle/``) is from the live reference on the
same streams.  The reference is imported, never copied.

Every fixture holds, per step t and env e:
  actions[t,e,4]   what ``action_sample()`` returned (manytor.py:215-217)
  fresh[t,e,x,3]   objectives the env was (re)initialised with immediately
                   before step t (NaN when it was not reset there)
  obs[t,e,3x], reward[t,e], done[t,e], joints[t,e,4,3], alive[t,e,x]
                   what ``step`` returned / left in the env (manytor.py:255-260)
plus obs0[e,3x] = ``reset(returnable=True)`` of the first reset.
"""
import os
import sys
import time

import numpy as np

REF = os.environ.get("MANYTOR_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, REF)

import manytor as ref  # noqa: E402  (the unmodified reference)

from oracle import OracleEnvs, fk as oracle_fk  # noqa: E402
from oracle.scalar_port import ScalarEnv, scalar_reset, scalar_step  # noqa: E402


def run_reference(seed, n_envs, x, steps, reset_on_done, use_multienv=False, epochs=1):
    """Drive the reference like test_single.py / test_multi.py (no render)."""
    np.random.seed(seed)
    if use_multienv:
        shape = use_multienv
        me = ref.Multienv(env_shape=shape, obj_number=x)
        envs = me.environment
        assert me.env_number == n_envs
    else:
        envs = [ref.Environment(obj_number=x, index=i) for i in range(n_envs)]
    T = steps * epochs
    rec = dict(
        actions=np.zeros((T, n_envs, 4)),
        fresh=np.full((T, n_envs, x, 3), np.nan),
        obs=np.zeros((T, n_envs, 3 * x)),
        reward=np.zeros((T, n_envs), dtype=np.int64),
        done=np.zeros((T, n_envs), dtype=bool),
        joints=np.zeros((T, n_envs, 4, 3)),
        alive=np.zeros((T, n_envs, x), dtype=bool),
        total_reward=np.zeros((T, n_envs)),
    )
    obs0 = np.zeros((n_envs, 3 * x))
    pending = np.zeros(n_envs, dtype=bool)
    for e, env in enumerate(envs):                  # Multienv.reset order, manytor.py:106-107
        obs0[e] = env.reset(returnable=True)
        pending[e] = True
    t = 0
    for ep in range(epochs):
        for _ in range(steps):
            if use_multienv:
                acts = me.action_sample()           # manytor.py:111-113
            for e, env in enumerate(envs):
                if pending[e]:
                    rec["fresh"][t, e] = env.points
                    pending[e] = False
                a = acts[e] if use_multienv else env.action_sample()
                rec["actions"][t, e] = a
                o, r, d = env.step(a)
                rec["obs"][t, e] = o
                rec["reward"][t, e] = r
                rec["done"][t, e] = d
                rec["joints"][t, e] = env.joints_coordinates
                rec["alive"][t, e] = env.alives
                rec["total_reward"][t, e] = env.total_reward
                if d and reset_on_done:             # test_single.py:20-21,32
                    env.reset()
                    pending[e] = True
            t += 1
        if ep + 1 < epochs:                         # test_multi.py:34
            for e, env in enumerate(envs):
                env.reset()
                pending[e] = True
    rec["obs0"] = obs0
    return rec


def replay_oracle(rec, x):
    T, n = rec["reward"].shape
    env = OracleEnvs(n, x)
    worst = dict(obs=0.0, joints=0.0, flags=0)
    for t in range(T):
        m = ~np.isnan(rec["fresh"][t, :, 0, 0])
        if m.any():
            o0 = env.reset(mask=m, points=np.nan_to_num(rec["fresh"][t]), returnable=True)
            if t == 0:
                worst["obs"] = max(worst["obs"], np.abs(o0 - rec["obs0"]).max())
        r = env.step(rec["actions"][t])
        worst["obs"] = max(worst["obs"], np.abs(r.obs - rec["obs"][t]).max())
        worst["joints"] = max(worst["joints"], np.abs(r.joints - rec["joints"][t]).max())
        worst["flags"] += int((r.reward != rec["reward"][t]).sum() + (r.done != rec["done"][t]).sum()
                              + (r.alive != rec["alive"][t]).sum())
        worst["flags"] += int((np.abs(env.total_reward - rec["total_reward"][t]) > 0).sum())
    return worst


def replay_scalar(rec, x, max_steps=40):
    T, n = rec["reward"].shape
    envs = [ScalarEnv(x) for _ in range(n)]
    bad, worst = 0, 0.0
    for t in range(min(T, max_steps)):
        for e in range(n):
            if not np.isnan(rec["fresh"][t, e, 0, 0]):
                scalar_reset(envs[e], rec["fresh"][t, e])
            o, r, d = scalar_step(envs[e], rec["actions"][t, e])
            worst = max(worst, np.abs(o - rec["obs"][t, e]).max())
            bad += int(r != rec["reward"][t, e]) + int(d != rec["done"][t, e])
    return dict(obs=worst, flags=bad)


def main():
    meta = dict(numpy=np.__version__, reference=REF, generated=time.strftime("%Y-%m-%d"))
    cases = {
        # BASELINE.json config 1 shape (test_single.py), shortened to 300 steps
        "single_x10_seed0": dict(seed=0, n_envs=1, x=10, steps=300, reset_on_done=True),
        # test_multi.py settings: Multienv((3,2), 7), 50 steps/epoch, reset between epochs
        "multi_3x2_x7_seed1": dict(seed=1, n_envs=6, x=7, steps=50, reset_on_done=False,
                                   use_multienv=(3, 2), epochs=2),
        # 32 envs, reset on done, long enough for several episodes to end
        "batch32_x10_seed2": dict(seed=2, n_envs=32, x=10, steps=150, reset_on_done=True),
        # few objectives -> frequent all-collected terminations and resets
        "batch16_x2_seed3": dict(seed=3, n_envs=16, x=2, steps=150, reset_on_done=True),
        # Multienv's own defaults: env_shape=(1, 2), obj_number=5 (manytor.py:77)
        "multi_default_1x2_x5_seed5": dict(seed=5, n_envs=2, x=5, steps=120, reset_on_done=False,
                                           use_multienv=(1, 2), epochs=2),
        # x=10 run long enough (mean episode ~550 steps) to see all-collected + refresh
        "batch8_x10_seed4": dict(seed=4, n_envs=8, x=10, steps=900, reset_on_done=True),
    }
    for name, kw in cases.items():
        t0 = time.time()
        rec = run_reference(**kw)
        dt = time.time() - t0
        w = replay_oracle(rec, kw["x"])
        s = replay_scalar(rec, kw["x"])
        n_steps = rec["reward"].size
        print(f"{name}: {n_steps} env-steps in {dt:.1f}s ({n_steps / dt:.0f}/s)  "
              f"oracle: obs {w['obs']:.2e} joints {w['joints']:.2e} flag-mismatch {w['flags']}  "
              f"scalar-port: obs {s['obs']:.2e} flag-mismatch {s['flags']}  "
              f"done-count {int(rec['done'].sum())} +1 {int((rec['reward'] == 1).sum())} "
              f"-1 {int((rec['reward'] == -1).sum())}")
        assert w["flags"] == 0 and w["obs"] < 1e-9 and w["joints"] < 1e-9, "oracle disagrees with reference"
        assert s["flags"] == 0 and s["obs"] < 1e-9
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=kw["x"], seed=kw["seed"],
                            ref_steps_per_s=n_steps / dt, **meta, **rec)

    # known-answer vectors, re-captured from the live reference
    env = ref.Environment(10)
    np.random.seed(0)
    env.reset()
    kat = dict(
        fk4_30_45_60_90=ref.fk(4, [30, 45, 60, 90]),
        fk3_30_45_60_90=ref.fk(3, [30, 45, 60, 90]),
        fk2_30_45_60_90=ref.fk(2, [30, 45, 60, 90]),
        zero_pose_joints=env.joints_coordinates.copy(),
        seed0_points=env.points.copy(),
        dh_sample=ref.dh(27.0, np.pi / 2, 0.5, 0.3),
        r_theta_sample=np.array(ref.r_theta([1.0, 2.0, 3.0], [4.0, -1.0, 0.5])),
    )
    assert np.abs(oracle_fk(4, [30, 45, 60, 90]) - kat["fk4_30_45_60_90"]).max() < 1e-12
    np.savez_compressed(os.path.join(HERE, "known_answers.npz"), **meta, **kat)
    print("known answers:", {k: np.asarray(v).round(6).tolist() for k, v in kat.items()
                             if np.asarray(v).size <= 4})


if __name__ == "__main__":
    main()
