"""Scalar, one-env-at-a-time port of the reference loop -- TEST INFRASTRUCTURE ONLY.

This is the "test_multi.py-style" CPU baseline: the reference itself cannot
travel to the GPU box (``/root/reference`` does not exist there), so the cost
structure of its loop -- one Python-level 4x4 build and one ``dot`` per DH row,
three ``fk`` calls per sub-pose, 25 sub-poses per step, a sequential ``for``
over the environments (manytor.py:115-122, 183-192, 35-53, 25-32) -- is
restated here in functional form and timed by ``bench.py`` next to the CUDA
path.  It is validated against the live reference by
``tests/golden/make_golden.py`` and against the vectorised oracle by
``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import math

import numpy as np

_ROWS = ((0.0, -math.pi / 2, 4.3, 0.0), (0.0, math.pi / 2, 0.0, 0.0),
         (0.0, -math.pi / 2, 24.3, 0.0), (27.0, math.pi / 2, 0.0, -math.pi / 2))


def _dh_row(a, alfa, d, theta):
    # manytor.py:25-32
    ct, st, ca, sa = np.cos(theta), np.sin(theta), np.cos(alfa), np.sin(alfa)
    return np.array([[ct, -st * ca, st * sa, a * ct],
                     [st, ct * ca, -ct * sa, a * st],
                     [0, sa, ca, d],
                     [0, 0, 0, 1]])


def _fk(mode, goals):
    # manytor.py:35-53
    t = [math.radians(g) for g in goals]
    m = np.eye(4)
    for i in range(mode):
        a, alfa, d, off = _ROWS[i]
        m = m.dot(_dh_row(a, alfa, d, t[i] + off))
    return m


def _joints(goals):
    # manytor.py:188-189
    return np.vstack((np.zeros(3), np.array([_fk(i, goals)[0:3, 3] for i in range(2, 5)])))


class ScalarEnv:
    """State of one environment (manytor.py:130-139, render fields dropped)."""

    __slots__ = ("x", "goals", "alive", "points", "joints", "total_reward")

    def __init__(self, x):
        self.x = x
        self.goals = np.zeros(4)
        self.alive = np.ones(x, dtype=bool)
        self.points = np.zeros((x, 3))
        self.joints = _joints(self.goals)
        self.total_reward = 0.0


def scalar_reset(env: ScalarEnv, points: np.ndarray) -> None:
    # manytor.py:219-241 with the objectives supplied by the caller
    env.goals = np.zeros(4)
    env.total_reward = 0.0
    env.alive = np.ones(env.x, dtype=bool)
    env.joints = _joints(env.goals)
    env.points = np.array(points, dtype=np.float64).reshape(env.x, 3)


def _obs(env: ScalarEnv):
    # manytor.py:141-153 and 17-22
    out = []
    e = env.joints[2]
    for p in range(env.x):
        if not env.alive[p]:
            out += [0.0, 0.0, 0.0]
            env.points[p] = 0.0
            continue
        d = [abs(e[i] - env.points[p, i]) for i in range(3)]
        h = math.sqrt(d[0] ** 2 + d[1] ** 2)
        out += [math.sqrt(h ** 2 + d[2] ** 2),
                math.degrees(math.atan2(d[0], d[1])),
                math.degrees(math.atan2(h, d[2]))]
    return np.array(out)


def _catch(env: ScalarEnv) -> bool:
    # manytor.py:155-173
    ee = env.joints[3]
    for p in range(env.x):
        if all(math.isclose(ee[a], env.points[p, a], abs_tol=8.0) for a in range(3)):
            env.alive[p] = False
    return not env.alive.any()


def scalar_step(env: ScalarEnv, action):
    # manytor.py:255-260 -> 175-213
    _obs(env)
    before = int(env.alive.sum())
    neg = False
    route = np.linspace(env.goals, action, num=25)
    for p in range(25):
        env.goals = route[p]
        env.joints = _joints(env.goals)
        if env.joints[2, 2] < 0 or env.joints[3, 2] < 0:
            neg = True
    obs2 = _obs(env)
    _catch(env)
    reward = 1 if before > int(env.alive.sum()) else 0
    if neg:
        reward = -1
    env.total_reward += reward
    return obs2, reward, _catch(env)


def multienv_loop(n_envs: int, x: int, steps: int, seed: int = 0):
    """``test_multi.py``-shaped loop (test_multi.py:11-23): build, reset, then
    ``steps`` x (sample + step) over ``n_envs`` sequential envs.  Returns the
    number of env-steps executed and the summed reward (a cheap checksum)."""
    rng = np.random.RandomState(seed)
    envs = [ScalarEnv(x) for _ in range(n_envs)]
    for e in envs:
        pts = []
        while len(pts) < x:                                         # manytor.py:228-239
            c = rng.uniform(-51.3, 51.3, size=3)
            if c[2] >= 0 and math.sqrt(c[0] ** 2 + c[1] ** 2 + c[2] ** 2) <= 51.3:
                pts.append(c)
        scalar_reset(e, np.array(pts))
    total = 0
    for _ in range(steps):
        for e in envs:
            a = rng.randint(-180, 180, size=4).astype(np.float64)   # manytor.py:215-217
            _, r, _ = scalar_step(e, a)
            total += r
    return n_envs * steps, total
