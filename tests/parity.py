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
