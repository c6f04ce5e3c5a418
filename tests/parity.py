"""Lock-step comparison of the CUDA path with the CPU oracle.

Parity bar (BASELINE.json north_star): joint / end-effector positions within
1e-5 relative (of the arm's reach) in fp32; reward, catch events (alive mask)
and done flags bit-exact EXCEPT where a distance lies within `THRESH_TOL` of the
catch or ground threshold -- those cases are counted and reported, and the
device state is re-synchronised to the oracle so the rest of the run stays
comparable.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

REL_POS_TOL = 1e-5          # of reach
THRESH_TOL = 1e-3           # |distance - threshold| below which a flag flip is "near-threshold"
DIST_ATOL_REL = 2e-5        # distance observation, relative to reach
ANGLE_FLOOR_DEG = 1e-3      # angle observation floor; plus the conditioning term below


@dataclass
class Report:
    env_steps: int = 0
    max_joint_err: float = 0.0
    max_dist_err: float = 0.0
    max_angle_excess: float = 0.0      # worst (error / allowed)
    near_threshold: int = 0            # reported, allowed
    hard_mismatch: int = 0             # must stay 0
    notes: list = field(default_factory=list)

    def ok(self) -> bool:
        return self.hard_mismatch == 0

    def summary(self) -> str:
        return (f"{self.env_steps} env-steps: joint err {self.max_joint_err:.2e}, dist err {self.max_dist_err:.2e}, "
                f"angle err/allowed {self.max_angle_excess:.2f}, near-threshold flips {self.near_threshold}, "
                f"hard mismatches {self.hard_mismatch}")


def reach_of(spec) -> float:
    tab = spec.table()
    return float(np.abs(tab[:, 0]).sum() + np.abs(tab[:, 2]).sum())


def compare_step(rep: Report, spec, ora, r, points_before, alive_before, dev_obs, dev_reward, dev_done,
                 dev_alive, dev_joints=None, scale=None):
    """Compare one step.  `r` is the oracle StepResult; returns the bool mask of
    envs whose flags differ (all of them near-threshold, or hard_mismatch grows)."""
    n, x = alive_before.shape
    reach = reach_of(spec)
    pos_tol = REL_POS_TOL * reach
    thr = THRESH_TOL * (reach / 55.6)
    rep.env_steps += n

    if dev_joints is not None:
        je = np.abs(dev_joints.astype(np.float64) - r.joints).max()
        rep.max_joint_err = max(rep.max_joint_err, float(je))
        if je > pos_tol:
            rep.hard_mismatch += 1
            rep.notes.append(f"joint error {je:.3e} > {pos_tol:.3e}")

    # flags
    rew_bad = dev_reward.astype(np.int64) != r.reward
    done_bad = (dev_done & 1).astype(bool) != r.done
    alive_bad = (dev_alive != r.alive).any(axis=1)
    bad = rew_bad | done_bad | alive_bad
    cm = np.where(alive_before, np.abs(r.catch_margin), np.inf).min(axis=1)
    near = (r.ground_margin < thr) | (cm < thr)
    rep.near_threshold += int((bad & near).sum())
    hard = bad & ~near
    if hard.any():
        rep.hard_mismatch += int(hard.sum())
        i = int(np.nonzero(hard)[0][0])
        rep.notes.append(f"env {i}: reward dev {dev_reward[i]} ora {r.reward[i]}, done dev {dev_done[i]} ora {r.done[i]}, "
                         f"ground margin {r.ground_margin[i]:.3e}, catch margin {cm[i]:.3e}")

    # observations (computed from alive_before, which is kept in sync)
    anchor = r.joints[:, ora._row(spec.obs_frame)]
    d = np.abs(anchor[:, None, :] - points_before)
    hxy = np.hypot(d[..., 0], d[..., 1])
    dist = np.sqrt(hxy ** 2 + d[..., 2] ** 2)
    o_dev = dev_obs.astype(np.float64).reshape(n, x, 3)
    o_ref = r.obs.reshape(n, x, 3)
    de = np.abs(o_dev[..., 0] - o_ref[..., 0])
    rep.max_dist_err = max(rep.max_dist_err, float(de.max()))
    dist_tol = DIST_ATOL_REL * reach + pos_tol
    # an angle moves by ~ position error / lever arm; allow that plus a floor
    allow_r = np.degrees(pos_tol / np.maximum(hxy, 1e-9)) + ANGLE_FLOOR_DEG
    allow_t = np.degrees(pos_tol / np.maximum(dist, 1e-9)) + ANGLE_FLOOR_DEG
    er = np.abs(o_dev[..., 1] - o_ref[..., 1]) / np.minimum(allow_r, 90.0)
    et = np.abs(o_dev[..., 2] - o_ref[..., 2]) / np.minimum(allow_t, 90.0)
    ex = max(float(er.max()), float(et.max()))
    rep.max_angle_excess = max(rep.max_angle_excess, ex)
    if de.max() > dist_tol or ex > 1.0:
        rep.hard_mismatch += 1
        rep.notes.append(f"obs mismatch: dist err {de.max():.3e} (tol {dist_tol:.3e}), angle err/allowed {ex:.2f}")
    dead = ~alive_before
    if dead.any() and np.abs(o_dev[dead]).max() != 0.0:
        rep.hard_mismatch += 1
        rep.notes.append("dead objective did not read as zeros")
    return bad


def alive_bits_to_matrix(words: np.ndarray, x: int) -> np.ndarray:
    w = words.astype(np.int64) & 0xFFFFFFFF
    return ((w[:, None] >> np.arange(x)[None, :]) & 1).astype(bool)


def lockstep(env, ora, spec, actions_fn, steps: int, on_done=None, rep: Report | None = None, joints=True):
    """Advance device `env` (BatchedEnvs) and oracle `ora` together for `steps`.

    actions_fn(t) -> (N, J) array.  on_done(t, done_mask) -> (N, X, 3) fresh
    points for the envs that ended (reset on both sides), or None for no reset.
    """
    rep = rep or Report()
    for t in range(steps):
        act = np.float32(np.asarray(actions_fn(t))).astype(np.float64)   # the device sees fp32 actions
        alive_before = ora.alive.copy()
        points_before = ora.points.copy()
        out = env.step(act.astype(np.float32), joints=joints)
        obs, rew, done = (o.cpu().numpy() for o in out[:3])
        jn = out[3].cpu().numpy() if joints else None
        r = ora.step(act)
        dev_alive = alive_bits_to_matrix(env.get_state()["alive"].cpu().numpy(), ora.x)
        bad = compare_step(rep, spec, ora, r, points_before, alive_before, obs, rew, done, dev_alive, jn)
        if bad.any():   # near-threshold flips: put the device back on the oracle's trajectory
            env.set_state(goals=ora.goals, alive=ora.alive, total_reward=ora.total_reward, mask=bad)
        if on_done is not None and r.done.any():
            fresh = on_done(t, r.done)
            if fresh is not None:
                ora.reset(mask=r.done, points=fresh)
                env.reset(mask=r.done)
                env.set_points(fresh, mask=r.done)
    return rep


class ChunkedOracle:
    """One OracleEnvs interface over several smaller ones (bounded temporaries at 2^18..2^20 envs).

    Exposes what `compare_step` / the lock-step loops use: n, x, spec, goals, alive, points,
    total_reward, ep_len, step(), reset(), _row()."""

    def __init__(self, n_envs: int, obj_number: int, spec, chunk: int = 1 << 16, **kw):
        from oracle import OracleEnvs
        self.n, self.x, self.spec = int(n_envs), int(obj_number), spec
        self.bounds = [(a, min(a + chunk, self.n)) for a in range(0, self.n, chunk)]
        self.parts = [OracleEnvs(b - a, obj_number, spec=spec, **kw) for a, b in self.bounds]

    def _cat(self, name):
        return np.concatenate([getattr(p, name) for p in self.parts], axis=0)

    goals = property(lambda self: self._cat("goals"))
    alive = property(lambda self: self._cat("alive"))
    points = property(lambda self: self._cat("points"))
    total_reward = property(lambda self: self._cat("total_reward"))
    ep_len = property(lambda self: self._cat("ep_len"))

    def _row(self, frame):
        return self.parts[0]._row(frame)

    def reset(self, mask=None, points=None):
        for (a, b), p in zip(self.bounds, self.parts):
            p.reset(mask=None if mask is None else mask[a:b], points=None if points is None else points[a:b])

    def set_alive(self, alive):
        for (a, b), p in zip(self.bounds, self.parts):
            p.alive = np.array(alive[a:b], dtype=bool)

    def step(self, action):
        from oracle import StepResult
        rs = [p.step(action[a:b]) for (a, b), p in zip(self.bounds, self.parts)]
        return StepResult(*[np.concatenate([getattr(r, f) for r in rs], axis=0) for f in StepResult._fields])


def lockstep_auto_reset(env, ora, spec, stream, actions_fn, steps: int, horizon: int, rep: Report | None = None,
                        check_points: bool = True):
    """Device auto-reset (in-kernel, objectives from the uploaded `stream` (E, N, X, 3)) against the oracle
    reset on the host with the same sets.  The device must already hold set 0 (env.reset() after
    set_objective_stream) and the oracle too.  Returns (report, expected statistics, reward slack, forked):
    a near-threshold catch/done flip changes WHICH envs reset, after which the two sides are no longer
    comparable -- the loop stops there and says so (`forked`)."""
    rep = rep or Report()
    n, x = ora.n, ora.x
    sets = stream.shape[0]
    episode = np.ones(n, dtype=np.int64)                   # resets so far
    stats = dict(env_steps=0, episodes=0, terminated=0, reward_sum=0, length_sum=0, catches=0)
    slack, forked = 0, False
    for t in range(steps):
        act = np.float32(np.asarray(actions_fn(t))).astype(np.float64)
        alive_before, points_before = ora.alive.copy(), ora.points.copy()
        obs, rew, done = (o.cpu().numpy() for o in env.step(act.astype(np.float32)))
        r = ora.step(act)
        ep_len, total = ora.ep_len, ora.total_reward
        trunc = (ep_len >= horizon) & ~r.done if horizon > 0 else np.zeros(n, dtype=bool)
        ended = r.done | trunc
        st = env.get_state()
        dev_alive_after = alive_bits_to_matrix(st["alive"].cpu().numpy(), x)
        dev_alive = np.where(ended[:, None], r.alive, dev_alive_after)       # ended envs were reset on the device
        bad = compare_step(rep, spec, ora, r, points_before, alive_before, obs, rew, done, dev_alive, None)
        stats["env_steps"] += n
        if not rep.ok():
            break
        if bad.any():
            fork = (((done & 1).astype(bool) != r.done) | (dev_alive != r.alive).any(axis=1)) & bad
            if fork.any():
                forked = True
                break
            slack += 2 * int((bad & ended).sum())
            if (bad & ~ended).any():
                env.set_state(total_reward=total, mask=bad & ~ended)
        assert np.array_equal((done & 2).astype(bool), trunc), "truncation flags differ"
        if ended.any():
            stats["episodes"] += int(ended.sum())
            stats["terminated"] += int(r.done.sum())
            stats["reward_sum"] += int(total[ended].sum())
            stats["length_sum"] += int(ep_len[ended].sum())
            stats["catches"] += int((x - r.alive[ended].sum(axis=1)).sum())
            fresh = stream[episode % sets, np.arange(n)]
            ora.reset(mask=ended, points=fresh)
            episode[ended] += 1
            assert np.all(dev_alive_after[ended]) and np.all(st["goals"].cpu().numpy()[ended] == 0)
            assert np.all(st["ep_len"].cpu().numpy()[ended] == 0) and np.all(st["total_reward"].cpu().numpy()[ended] == 0)
            if check_points:
                np.testing.assert_array_equal(env.get_points(zero_dead=False).cpu().numpy()[ended],
                                              np.float32(fresh[ended]))
    return rep, stats, slack, forked


def half_ball_points(rs, shape, radius=51.3):
    """i.i.d. uniform points in the upper half ball (the law of manytor.py:228-239), vectorised."""
    v = rs.normal(size=tuple(shape) + (3,))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    v[..., 2] = np.abs(v[..., 2])
    return v * (radius * rs.uniform(0, 1, size=tuple(shape) + (1,)) ** (1 / 3))
