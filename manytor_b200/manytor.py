"""Drop-in for the reference's ``manytor`` module, backed by the CUDA step loop.

Same names, arguments and return conventions as the reference
(``import manytor_b200.manytor as tor`` in place of ``import manytor as tor``):

  Environment(obj_number, index)   reset / step / action_sample / render   manytor.py:125-283
  Multienv(env_shape, obj_number)  reset / step / action_sample / render   manytor.py:72-122
  fk(mode, goals), dh(a, alfa, d, theta), r_theta(v1, v2)                   manytor.py:17-53
  HOST, PORT                                                                manytor.py:9-10

The sequential per-env loop of ``Multienv`` (manytor.py:117-118) is replaced by
one fused kernel launch over all N = rows x cols environments.  Differences a
caller can see:
  * observations are computed in fp32 (returned as float64 arrays, like the
    reference's dtype);
  * objectives are drawn by the on-device sampler (Philox) instead of the global
    ``np.random`` stream -- same distribution, different numbers; ``set_points``
    uploads reference-generated objectives for parity runs;
  * ``Multienv.step`` returns lists like the reference by default; pass
    ``as_lists=False`` to get (N, 3X)/(N,)/(N,) numpy arrays over pinned host
    buffers without the per-env Python objects;
  * rendering is off the hot path: frames are produced from a copy-back of the
    rendered env(s) only (``render_envs``), over the reference's UDP protocol;
  * ``trajectory`` (manytor.py:135,190,223: the terminal's position at every one of the
    25 sub-poses of every step since the last reset) is materialised when it is READ, from
    the logged (pose before, action) pairs, by one batched kinematics launch -- the step
    itself only appends the pair.  ``Multienv`` logs it for up to 4096 envs by default
    (``track_trajectory``);
  * with ``horizon`` > 0, ``done`` is also True on the step that reaches the horizon
    (truncation; the reference's driver scripts cut episodes at ``max_steps`` the same way).
There is no CPU fallback: without the CUDA library and a B200 these classes raise.
"""
from __future__ import annotations

import json
import math
import socket
import threading
import time
from subprocess import call
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .core import ArmSpec, BatchedEnvs, REFERENCE_ARM, make_config

HOST = "localhost"   # manytor.py:9
PORT = 5001          # manytor.py:10
_SUBSTEPS = 25       # manytor.py:178


# ----------------------------------------------------------------------------
# module-level helpers of the reference, computed on the GPU
# ----------------------------------------------------------------------------
def _dev():
    if not torch.cuda.is_available():
        raise _lib.MantorLibraryError("no CUDA device: manytor_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream_ptr(device):
    import ctypes as C
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def fk(mode, goals, arm: ArmSpec = REFERENCE_ARM):
    """Forward kinematics -- manytor.py:35-53.  ``goals`` in degrees; returns the
    4x4 transform of frame ``mode`` (or (M, 4, 4) for (M, J) goals)."""
    import ctypes as C
    device = _dev()
    g = torch.as_tensor(np.asarray(goals, dtype=np.float32), device=device).reshape(-1, arm.n_joints).contiguous()
    out = torch.empty((g.shape[0], 16), dtype=torch.float32, device=device)
    cfg = make_config(1, 1, arm, device.index)
    _lib.check(_lib.load().mt_fk(C.byref(cfg), int(mode), C.c_void_p(g.data_ptr()), C.c_void_p(out.data_ptr()),
                                 g.shape[0], _stream_ptr(device)))
    m = out.cpu().numpy().astype(np.float64).reshape(-1, 4, 4)
    return m[0] if np.ndim(goals) == 1 else m


def dh(a, alfa, d, theta):
    """One DH transform -- manytor.py:25-32 (theta in radians); scalars or (M,) arrays."""
    import ctypes as C
    device = _dev()
    p = np.stack(np.broadcast_arrays(*[np.asarray(v, dtype=np.float32) for v in (a, alfa, d, theta)]), axis=-1)
    scalar = p.ndim == 1
    t = torch.as_tensor(p.reshape(-1, 4), device=device).contiguous()
    out = torch.empty((t.shape[0], 16), dtype=torch.float32, device=device)
    _lib.check(_lib.load().mt_dh(C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()), t.shape[0], _stream_ptr(device)))
    m = out.cpu().numpy().astype(np.float64).reshape(-1, 4, 4)
    return m[0] if scalar else m


def r_theta(v1, v2):
    """Bearing angles (degrees) of |v1 - v2| -- manytor.py:17-22."""
    import ctypes as C
    device = _dev()
    a = torch.as_tensor(np.asarray(v1, dtype=np.float32), device=device).reshape(-1, 3).contiguous()
    b = torch.as_tensor(np.asarray(v2, dtype=np.float32), device=device).reshape(-1, 3).contiguous()
    out = torch.empty((a.shape[0], 2), dtype=torch.float32, device=device)
    _lib.check(_lib.load().mt_r_theta(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()),
                                      a.shape[0], _stream_ptr(device)))
    o = out.cpu().numpy().astype(np.float64)
    return (float(o[0, 0]), float(o[0, 1])) if np.ndim(v1) == 1 else (o[:, 0], o[:, 1])


def plot_vispy():
    call(["python3", "ManyTor/plotting.py"])   # manytor.py:13-14 (the reference's viewer, unchanged)


class StoppableThread(threading.Thread):
    """manytor.py:56-69."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._stop_event = threading.Event()

    def stop(self):
        self._stop_event.set()

    def stopped(self):
        return self._stop_event.is_set()


# ----------------------------------------------------------------------------
# render shim: one-env copy-back -> the reference's UDP/JSON frames
# ----------------------------------------------------------------------------
class _Renderer:
    """Feeds plotting.py's protocol (plotting.py:27-87) from copied-back state.

    Per step and rendered env: 25 frames [index(3), joints(4x3), points(Xx3),
    trajectory tail(3)] like manytor.py:196-201; the 25 sub-pose joint positions
    come from one small ``mt_joints`` launch over the interpolated route."""

    def __init__(self, envs: BatchedEnvs, frame_sleep: float = 0.006):
        self.envs = envs
        self.udp = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        self.dest = (HOST, PORT)
        self.frame_sleep = frame_sleep
        self.first = {}

    def send(self, payload):
        self.udp.sendto(json.dumps(payload).encode(), self.dest)

    def frames(self, index: int, goals_before: np.ndarray, action: np.ndarray, points: np.ndarray):
        route = np.linspace(np.asarray(goals_before, dtype=np.float64), np.asarray(action, dtype=np.float64),
                            num=_SUBSTEPS)                                   # manytor.py:182
        joints = self.envs.joints_of(route.astype(np.float32)).cpu().numpy().astype(np.float64)
        for p in range(_SUBSTEPS):
            flag = 1 if self.first.get(index, True) else 0                    # manytor.py:196
            self.first[index] = False
            stacked = np.vstack(([index, np.nan, flag], joints[p], points, joints[p][-1]))
            self.send(np.squeeze(stacked.reshape(1, -1)).tolist())            # manytor.py:197-201
            if self.frame_sleep:
                time.sleep(self.frame_sleep)                                  # manytor.py:202

    def close(self):
        self.udp.close()


def _trajectory_rows(envs: BatchedEnvs, pairs) -> np.ndarray:
    """Terminal positions over the 25 interpolated poses of each logged (pose before, action) pair
    (manytor.py:182-190), (25 * len(pairs), 3) float64, by one batched ``mt_joints`` launch."""
    if not pairs:
        return np.zeros((0, 3))
    before = np.stack([p[0] for p in pairs]).astype(np.float64)
    action = np.stack([p[1] for p in pairs]).astype(np.float64)
    k = np.arange(_SUBSTEPS, dtype=np.float64)[None, :, None] / (_SUBSTEPS - 1)
    route = before[:, None, :] + k * (action - before)[:, None, :]
    route[:, -1, :] = action                                                   # linspace pins the endpoint
    joints = envs.joints_of(route.reshape(-1, envs.j).astype(np.float32)).cpu().numpy().astype(np.float64)
    return joints[:, -1, :]


class _TrajectoryLog:
    """``trajectory`` of one env: the nominal start row of manytor.py:223 plus 25 rows per step."""

    def __init__(self):
        self.clear()

    def clear(self):
        self.pairs, self.rows, self.done_pairs = [], np.array([[0.0, 0.0, 51.3]]), 0

    def append(self, before, action):
        self.pairs.append((np.array(before, dtype=np.float64), np.array(action, dtype=np.float64)))

    def materialise(self, envs: BatchedEnvs) -> np.ndarray:
        if self.done_pairs < len(self.pairs):
            self.rows = np.vstack((self.rows, _trajectory_rows(envs, self.pairs[self.done_pairs:])))
            self.done_pairs = len(self.pairs)
        return self.rows if len(self.pairs) else self.rows[0]                  # manytor.py:135: a (3,) vector before the first step


def _host_action_sample(cfg, n_envs):
    """[[np.random.randint(low, high) for _ in range(J)] for each env] -- manytor.py:215-217 and :111-113; one
    vectorised call consumes the legacy np.random stream exactly like the reference's scalar calls."""
    a = np.random.randint(int(cfg.action_low), int(cfg.action_high), size=(n_envs, int(cfg.n_joints)))
    return [[int(v) for v in row] for row in a]


class _EnvView:
    """``multienv.environment[i]``: one env of the batch with the reference's per-env surface
    (manytor.py:125-260; test_multi.py:32 reads ``total_reward``).  Attribute reads copy that env's
    state back (``mt_fetch_env``); ``step`` advances ONLY this env -- its state goes through a
    one-env handle and back, so it is for occasional use, the batch is what ``Multienv.step`` is for."""

    def __init__(self, owner: "Multienv", index: int):
        self._owner, self.id = owner, index
        self.obj_number = owner.obj_number

    def _fetch(self):
        return self._owner._envs.fetch_env(self.id)

    @property
    def total_reward(self):
        return self._fetch()["total_reward"]

    @property
    def goals(self):
        return self._fetch()["goals"].astype(np.float64)

    @property
    def alives(self):
        return self._fetch()["alives"]

    @property
    def points(self):
        return self._fetch()["points"].astype(np.float64)

    @property
    def joints_coordinates(self):
        return self._fetch()["joints_coordinates"].astype(np.float64)

    @property
    def rendering(self):
        return self._owner.rendering and self.id in self._owner.render_envs

    @property
    def trajectory(self):
        return self._owner._trajectory_of(self.id)

    def _one_hot(self):
        m = np.zeros(self._owner.env_number, dtype=bool)
        m[self.id] = True
        return m

    def get_observations(self):
        """manytor.py:141-153."""
        return self._owner._envs.observe().cpu().numpy().astype(np.float64)[self.id]

    def is_done(self):
        """manytor.py:155-173 on the current state."""
        return not bool(self.alives.any())

    def action_sample(self):
        """manytor.py:215-217: J draws from the process-global np.random stream, like the reference."""
        return _host_action_sample(self._owner._envs.cfg, 1)[0]

    def reset(self, returnable=False):
        """manytor.py:219-253 for this env only."""
        self._owner._envs.reset(mask=self._one_hot())
        self._owner._trajectory_reset(self.id)
        return self.get_observations() if returnable else None

    def step(self, action):
        """manytor.py:255-260 for this env only -> (obs2, reward, done)."""
        return self._owner._step_one(self.id, action)

    def render(self, stop_render=False, multienv=True):
        """manytor.py:262-283: add this env to / remove it from the envs streamed to the viewer."""
        envs = set(self._owner.render_envs)
        (envs.discard if stop_render else envs.add)(self.id)
        self._owner.render_envs = tuple(sorted(envs))


class _EnvList(Sequence):
    def __init__(self, owner):
        self._owner = owner

    def __len__(self):
        return self._owner.env_number

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [_EnvView(self._owner, k) for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return _EnvView(self._owner, i)


class Multienv:
    """Multienv(env_shape, obj_number) -- manytor.py:72-122, N = rows x cols envs in HBM."""

    def __init__(self, env_shape=(1, 2), obj_number=5, as_lists: bool = True, device=None, seed: int = 0,
                 arm: ArmSpec = REFERENCE_ARM, horizon: int = 0, auto_reset: bool = False,
                 terminate_on_ground: bool = False, render_envs: Sequence[int] = (0,),
                 track_trajectory: Optional[bool] = None, **kw):
        self.env_shape = env_shape
        self.env_number = env_shape[0] * env_shape[1]
        self.obj_number = obj_number
        self.rendering = False
        self.as_lists = as_lists
        self.render_envs = tuple(render_envs)
        self._kw = dict(arm=arm, device=device, seed=seed, horizon=horizon, terminate_on_ground=terminate_on_ground, **kw)
        self._envs = BatchedEnvs(self.env_number, obj_number, auto_reset=auto_reset, **self._kw)
        self._horizon, self._auto_reset = int(horizon), bool(auto_reset)
        self.environment = _EnvList(self)
        self._renderer: Optional[_Renderer] = None
        self._solo: Optional[BatchedEnvs] = None
        # trajectory log: per step the (N, J) actions; the pose before a step is the previous action
        # (zeros after a reset).  In-kernel auto-reset moves poses without the host seeing it: no log then.
        if track_trajectory is None:
            track_trajectory = self.env_number <= 4096
        self._track = bool(track_trajectory) and not auto_reset
        self._traj_actions: List[np.ndarray] = []
        self._traj_start = np.zeros(self.env_number, dtype=np.int64)           # first logged step of each env's episode
        self._traj_solo = {}                                                   # env -> _TrajectoryLog after per-env calls

    # the batched engine, for callers that want device tensors
    @property
    def batched(self) -> BatchedEnvs:
        return self._envs

    def reset(self, returnable=False):
        """manytor.py:106-109."""
        obs = self._envs.reset(returnable=returnable)
        self._traj_actions, self._traj_solo = [], {}
        self._traj_start[:] = 0
        if self._renderer is not None:
            self._renderer.send([float("nan"), float("nan"), 4])               # manytor.py:246-249
        if returnable:
            o = obs.cpu().numpy().astype(np.float64)
            return list(o) if self.as_lists else o
        return None

    # -- per-env surface behind ``environment[i]`` ------------------------------
    def _trajectory_of(self, i: int) -> np.ndarray:
        if i in self._traj_solo:
            return self._traj_solo[i].materialise(self._envs)
        if not self._track:
            return np.array([0.0, 0.0, 51.3])
        log = _TrajectoryLog()
        before = np.zeros(self._envs.j)
        for a in self._traj_actions[int(self._traj_start[i]):]:
            log.append(before, a[i])
            before = a[i]
        return log.materialise(self._envs)

    def _solo_log(self, i: int) -> "_TrajectoryLog":
        """Switch env i to its own log (it no longer moves in lock step with the logged batch actions)."""
        if i not in self._traj_solo:
            log = _TrajectoryLog()
            if self._track:
                before = np.zeros(self._envs.j)
                for a in self._traj_actions[int(self._traj_start[i]):]:
                    log.append(before, a[i])
                    before = a[i]
            self._traj_solo[i] = log
        return self._traj_solo[i]

    def _trajectory_reset(self, i: int):
        self._traj_solo.pop(i, None)
        self._traj_start[i] = len(self._traj_actions)

    def _step_one(self, i: int, action):
        if self._solo is None:
            self._solo = BatchedEnvs(1, self.obj_number, auto_reset=False, **{**self._kw, "env_id_base": 0})
            self._solo.reset()
        st = self._envs.get_state()
        one = {k: v[i:i + 1] for k, v in st.items()}
        self._solo.set_points(self._envs.get_points(zero_dead=False)[i:i + 1])
        self._solo.set_state(goals=one["goals"], alive=one["alive"], total_reward=one["total_reward"], ep_len=one["ep_len"])
        before = one["goals"].cpu().numpy()[0].astype(np.float64)
        act = np.asarray(action, dtype=np.float32).reshape(1, self._envs.j)
        obs, rew, done = self._solo.step(act)
        new = self._solo.get_state()
        mask = np.zeros(self.env_number, dtype=bool)
        mask[i] = True
        full = {k: st[k].clone() for k in st}
        for k in full:
            full[k][i] = new[k][0]
        self._envs.set_state(goals=full["goals"], alive=full["alive"], total_reward=full["total_reward"],
                             ep_len=full["ep_len"], mask=mask)
        self._solo_log(i).append(before, act[0])
        d = int(done.cpu().numpy()[0])
        return (obs.cpu().numpy().astype(np.float64)[0], int(rew.cpu().numpy()[0]),
                bool(d) if self._horizon > 0 else bool(d & 1))

    def action_sample(self):
        """manytor.py:111-113: per env J integers in [-180, 180).  In list mode they come from the process-global
        np.random stream in the reference's order (env 0's J draws, then env 1's, ...); in array mode from the
        device sampler (`mt_sample_actions`, the stream `rollout_random` uses), read back as an (N, J) int64 array."""
        if self.as_lists:
            return _host_action_sample(self._envs.cfg, self.env_number)
        return self._envs.sample_actions().cpu().numpy().astype(np.int64)

    def step(self, action):
        """manytor.py:115-122 -> (obs2, reward, done) per env."""
        if self.as_lists or self._renderer is not None:
            act = np.asarray(action, dtype=np.float32).reshape(self.env_number, self._envs.j)
        else:
            act = action
        before = None
        if self._renderer is not None:
            before = {i: self._envs.fetch_env(i) for i in self.render_envs if i < self.env_number}
        obs, rew, done = self._envs.step_host(act)
        if self._track:
            logged = np.array(act, dtype=np.float32).reshape(self.env_number, self._envs.j)   # a copy: the caller may reuse `act`
            self._traj_actions.append(logged)
            for i, log in self._traj_solo.items():                             # envs stepped on their own keep their own log
                prev = log.pairs[-1][1] if log.pairs else np.zeros(self._envs.j)
                log.append(prev, logged[i])
        if self._renderer is not None:
            for i, st in before.items():
                self._renderer.frames(i, st["goals"], act[i], st["points"])
        if self.as_lists:
            o = obs.astype(np.float64)
            # bit0 = all objectives collected (manytor.py:170-171); bit1 = horizon reached (only with horizon > 0)
            return list(o), [int(r) for r in rew], [bool(d) if self._horizon > 0 else bool(d & 1) for d in done]
        return obs, rew, done

    def render(self, stop_render=False):
        """manytor.py:84-104; only ``render_envs`` are streamed to the viewer."""
        if not stop_render:
            self.processThread = StoppableThread(target=plot_vispy)
            self.processThread.start()
            self._renderer = _Renderer(self._envs)
            self.rendering = True
            time.sleep(0.75)
            self._renderer.send([self.env_number, self.obj_number, 3, list(self.env_shape)])  # manytor.py:94-97
        else:
            self.rendering = False
            if self._renderer is not None:
                self._renderer.send([float("nan"), float("nan"), 2])           # JSON, unlike manytor.py:100 (C10)
                time.sleep(0.5)
                self._renderer.close()
                self._renderer = None
            if getattr(self, "processThread", None) is not None:
                self.processThread.stop()


class Environment:
    """Environment(obj_number, index) -- manytor.py:125-283, a batch of one."""

    def __init__(self, obj_number=10, index=0, device=None, seed: int = 0, arm: ArmSpec = REFERENCE_ARM, **kw):
        self.id = index
        self.obj_number = obj_number
        self.rendering = False
        self._traj = _TrajectoryLog()                                          # manytor.py:135
        self._pose = np.zeros(arm.n_joints)                                    # host mirror of `goals` (pose before the next step)
        self._horizon = int(kw.get("horizon", 0))
        self._envs = BatchedEnvs(1, obj_number, arm=arm, device=device, seed=seed, env_id_base=index, **kw)
        self._renderer: Optional[_Renderer] = None

    @property
    def trajectory(self):
        """manytor.py:135,190,223: terminal position at every sub-pose of every step since reset()."""
        return self._traj.materialise(self._envs)

    # state attributes the reference exposes (manytor.py:131-139)
    @property
    def goals(self):
        return self._envs.fetch_env(0)["goals"].astype(np.float64)

    @property
    def alives(self):
        return self._envs.fetch_env(0)["alives"]

    @property
    def points(self):
        return self._envs.fetch_env(0)["points"].astype(np.float64)

    @points.setter
    def points(self, value):
        self._envs.set_points(np.asarray(value, dtype=np.float32).reshape(1, self.obj_number, 3))

    @property
    def joints_coordinates(self):
        return self._envs.fetch_env(0)["joints_coordinates"].astype(np.float64)

    @property
    def total_reward(self):
        return self._envs.fetch_env(0)["total_reward"]

    def get_observations(self):
        """manytor.py:141-153."""
        return self._envs.observe().cpu().numpy().astype(np.float64)[0]

    def is_done(self):
        """manytor.py:155-173 evaluated on the current state (no mutation needed:
        catches are applied inside ``step``)."""
        return not bool(self.alives.any())

    def action_sample(self):
        """manytor.py:215-217: J draws of np.random.randint(-180, 180) from the process-global stream, so a seeded
        caller sees the reference's actions (no device round trip: a batch of one is pure latency)."""
        return _host_action_sample(self._envs.cfg, 1)[0]

    def reset(self, returnable=False):
        """manytor.py:219-253."""
        self._traj.clear()
        self._pose = np.zeros(self._envs.j)
        obs = self._envs.reset(returnable=returnable)
        if self._renderer is not None:
            self._renderer.send([float("nan"), float("nan"), 4])
        if returnable:
            return obs.cpu().numpy().astype(np.float64)[0]
        return None

    def step(self, action):
        """manytor.py:255-260 -> (obs2 (3X,) float64, reward int, done bool)."""
        act = np.asarray(action, dtype=np.float32).reshape(1, self._envs.j)
        before = self._envs.fetch_env(0) if self._renderer is not None else None
        # host-buffer step: upload, kernel and read-back in ONE synchronous call (a batch of one is pure latency)
        obs, rew, done = self._envs.step_host(act)
        obs = obs.astype(np.float64)[0]
        d = int(done[0])
        self._traj.append(self._pose, act[0])                                  # manytor.py:190, materialised on read
        self._pose = act[0].astype(np.float64)
        if d and self._envs.cfg.auto_reset:                                    # the kernel has already reset this env
            self._traj.clear()
            self._pose = np.zeros(self._envs.j)
        if before is not None:
            self._renderer.frames(self.id, before["goals"], act[0], before["points"])
        return obs, int(rew[0]), bool(d) if self._horizon > 0 else bool(d & 1)

    def render(self, stop_render=False, multienv=False):
        """manytor.py:262-283."""
        if not stop_render:
            if not multienv:
                self.processThread = StoppableThread(target=plot_vispy)
                self.processThread.start()
            self.rendering = True
            self._renderer = _Renderer(self._envs)
            if not multienv:
                time.sleep(0.75)
                self._renderer.send([1, self.obj_number, 3])                   # manytor.py:271-274
        else:
            self.rendering = False
            if self._renderer is not None:
                self._renderer.send([float("nan"), float("nan"), 2])           # manytor.py:277-279
                time.sleep(0.5)
                self._renderer.close()
                self._renderer = None
            if not multienv and getattr(self, "processThread", None) is not None:
                self.processThread.stop()
