"""Device-level API: N lock-step ManyTor environments resident on one B200.

`BatchedEnvs` owns one C-ABI handle (include/manytor_b200.h) and exposes the
reference's operations (reset / step / action_sample / get_observations,
manytor.py:141-260) on torch CUDA tensors.  torch is plumbing only: device
memory, streams and (in distributed.py) the NCCL process group.  All compute
runs in libmanytor_b200.so; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import MantorLibraryError, MtConfig, MtStats, STATS_FIELDS


@dataclass(frozen=True)
class ArmSpec:
    """A DH arm: rows (a, alpha, d, theta_offset) in radians plus the frame
    selectors the reference hard-codes (manytor.py:42-48, 143, 162, 191)."""

    dh: tuple
    obs_frame: int
    ground_frames: tuple
    catch_frame: int
    radius: float = 51.3
    catch_tol: float = 8.0

    @property
    def n_joints(self) -> int:
        return len(self.dh)


REFERENCE_ARM = ArmSpec(
    dh=((0.0, -math.pi / 2, 4.3, 0.0), (0.0, math.pi / 2, 0.0, 0.0),
        (0.0, -math.pi / 2, 24.3, 0.0), (27.0, math.pi / 2, 0.0, -math.pi / 2)),
    obs_frame=3, ground_frames=(3, 4), catch_frame=4)

# BASELINE.json config 5 (UR5-style 6-DOF chain, metres); selectors per SURVEY.md 8(a-FK)
UR5_ARM = ArmSpec(
    dh=((0.0, math.pi / 2, 0.089159, 0.0), (-0.425, 0.0, 0.0, 0.0), (-0.39225, 0.0, 0.0, 0.0),
        (0.0, math.pi / 2, 0.10915, 0.0), (0.0, -math.pi / 2, 0.09465, 0.0), (0.0, 0.0, 0.0823, 0.0)),
    obs_frame=5, ground_frames=(5, 6), catch_frame=6,
    radius=0.81725, catch_tol=0.81725 * 8.0 / 51.3)


def make_config(n_envs: int, obj_number: int, arm: ArmSpec, device: int, env_id_base: int = 0,
                horizon: int = 0, auto_reset: bool = False, terminate_on_ground: bool = False,
                obs_after_reset: bool = False, seed: int = 0, fk_mode: int = 0, substeps: int = 25,
                action_low: int = -180, action_high: int = 180) -> MtConfig:
    cfg = _lib.default_config()
    cfg.device = int(device)
    cfg.n_envs = int(n_envs)
    cfg.env_id_base = int(env_id_base)
    cfg.n_joints = arm.n_joints
    cfg.n_obj = int(obj_number)
    for i, row in enumerate(arm.dh):
        for k in range(4):
            cfg.dh[i][k] = float(row[k])
    cfg.obs_frame = arm.obs_frame
    cfg.ground_frame_a, cfg.ground_frame_b = arm.ground_frames
    cfg.catch_frame = arm.catch_frame
    cfg.radius = arm.radius
    cfg.catch_tol = arm.catch_tol
    cfg.substeps = int(substeps)
    cfg.horizon = int(horizon)
    cfg.terminate_on_ground = int(bool(terminate_on_ground))
    cfg.auto_reset = int(bool(auto_reset))
    cfg.obs_after_reset = int(bool(obs_after_reset))
    cfg.fk_mode = int(fk_mode)
    cfg.action_low = int(action_low)
    cfg.action_high = int(action_high)
    cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return cfg


def _device_index(device) -> int:
    if device is None:
        if not torch.cuda.is_available():
            raise MantorLibraryError("no CUDA device: manytor_b200 needs a B200 (there is no CPU fallback)")
        return torch.cuda.current_device()
    if isinstance(device, int):
        return device
    d = torch.device(device)
    if d.type != "cuda":
        raise MantorLibraryError(f"device {d} is not a CUDA device (there is no CPU fallback)")
    return d.index if d.index is not None else torch.cuda.current_device()


class _PinnedBlock:
    """Owner of one page-locked allocation (mt_host_alloc).  The numpy views handed to callers keep
    it alive through their base chain (view -> ctypes buffer -> block), so the memory is released
    only when the last view is gone -- not when the BatchedEnvs that allocated it is collected."""

    def __init__(self, nbytes: int, alloc=None, free=None):
        if (alloc is None) != (free is None):
            raise ValueError("pass both `alloc` and `free` (tests) or neither (mt_host_alloc / mt_host_free)")
        ptr = C.c_void_p()
        if alloc is None:
            lib = _lib.load()
            self._free = lib.mt_host_free
            _lib.check(lib.mt_host_alloc(C.byref(ptr), max(int(nbytes), 1)))
        else:
            self._free = free
            ptr = C.c_void_p(alloc(max(int(nbytes), 1)))
        self.ptr = ptr.value
        self._finalizer = weakref.finalize(self, self._free, C.c_void_p(self.ptr))


class PinnedArray:
    """A numpy array over page-locked host memory (mt_host_alloc)."""

    def __init__(self, shape, dtype, alloc=None, free=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        count = int(np.prod(self.shape))
        nbytes = count * self.dtype.itemsize
        self.block = _PinnedBlock(nbytes, alloc, free)
        buf = (C.c_char * max(nbytes, 1)).from_address(self.block.ptr)
        buf._mt_block = self.block                     # the buffer object is the base of every view below
        self.array = np.frombuffer(buf, dtype=self.dtype, count=count).reshape(self.shape)

    @property
    def ptr(self) -> int:
        return self.block.ptr


class BatchedEnvs:
    """N environments in HBM advanced by one fused kernel per step."""

    def __init__(self, n_envs: int, obj_number: int = 10, arm: ArmSpec = REFERENCE_ARM, device=None,
                 env_id_base: int = 0, horizon: int = 0, auto_reset: bool = False,
                 terminate_on_ground: bool = False, obs_after_reset: bool = False, seed: int = 0,
                 fk_mode: int = 0, substeps: int = 25, action_low: int = -180, action_high: int = 180,
                 lib_path: Optional[str] = None):
        self._lib = _lib.load(lib_path)
        self._h = C.c_void_p()
        dev = _device_index(device)
        self.device = torch.device("cuda", dev)
        self.arm = arm
        self.n, self.x, self.j = int(n_envs), int(obj_number), arm.n_joints
        self.cfg = make_config(n_envs, obj_number, arm, dev, env_id_base, horizon, auto_reset,
                               terminate_on_ground, obs_after_reset, seed, fk_mode, substeps,
                               action_low, action_high)
        _lib.check(self._lib.mt_create(C.byref(self.cfg), C.byref(self._h)))
        self._obs = self._reward = self._done = self._joints = self._actions = None
        self._stream_keepalive = None
        self._pinned = {}
        self._pinned_ptr = {}                              # id(array) -> address, for the arrays in _pinned (kept alive there)
        self._host_io = {}                                 # write_obs -> step_host's buffers and their addresses

    # -- plumbing -------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.mt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def _as_dev(self, a, dtype, shape=None):
        t = torch.as_tensor(a)
        t = t.to(device=self.device, dtype=dtype).contiguous()
        if shape is not None:
            t = t.reshape(shape)
        return t

    @staticmethod
    def _p(t: Optional[torch.Tensor]):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def _mask(self, mask):
        if mask is None:
            return None
        return self._as_dev(mask, torch.bool, (self.n,)).view(torch.uint8)

    def _out_buffers(self, want_obs=True, want_joints=False):
        if self._reward is None:
            self._reward = self._empty((self.n,), torch.float32)
            self._done = self._empty((self.n,), torch.uint8)
        if want_obs and self._obs is None:
            self._obs = self._empty((self.n, 3 * self.x), torch.float32)
        if want_joints and self._joints is None:
            self._joints = self._empty((self.n, self.j, 3), torch.float32)

    # -- reference operations --------------------------------------------------
    def reset(self, mask=None, returnable: bool = False):
        """Environment.reset / Multienv.reset (manytor.py:219-253, 106-109)."""
        m = self._mask(mask)
        _lib.check(self._lib.mt_reset(self._h, self._p(m), self._stream()))
        if returnable:
            # the same storage `step` writes its observations to: a policy that holds on to the tensor
            # returned here keeps seeing fresh observations (examples/policy_loop.py)
            self._out_buffers()
            return self.observe(out=self._obs)
        return None

    def observe(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Environment.get_observations (manytor.py:141-153) for every env -> (N, 3X)."""
        out = out if out is not None else self._empty((self.n, 3 * self.x), torch.float32)
        _lib.check(self._lib.mt_observe(self._h, self._p(out), self._stream()))
        return out

    def step(self, actions, joints: bool = False, write_obs: bool = True):
        """Environment.step / Multienv.step (manytor.py:255-260, 115-122).

        actions: (N, J) degrees.  Returns (obs (N,3X) f32, reward (N,) f32,
        done (N,) u8 [, joints (N,J,3)]) -- views of buffers this object reuses
        on the next step.
        """
        if (isinstance(actions, torch.Tensor) and actions.is_cuda and actions.dtype == torch.float32
                and actions.is_contiguous() and actions.numel() == self.n * self.j):
            a = actions                                    # fast path: already the device layout
        else:
            a = self._as_dev(actions, torch.float32, (self.n, self.j))
        self._out_buffers(write_obs, joints)
        _lib.check(self._lib.mt_step(self._h, self._p(a), self._p(self._obs if write_obs else None),
                                     self._p(self._reward), self._p(self._done),
                                     self._p(self._joints if joints else None), self._stream()))
        self._actions = a  # keep alive until the stream has consumed it
        out = (self._obs if write_obs else None, self._reward, self._done)
        return out + (self._joints,) if joints else out

    def sample_actions(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Multienv.action_sample (manytor.py:111-113, 215-217): integer degrees, (N, J) f32."""
        out = out if out is not None else self._empty((self.n, self.j), torch.float32)
        _lib.check(self._lib.mt_sample_actions(self._h, self._p(out), self._stream()))
        return out

    def rollout_random(self, n_steps: int, write_obs: bool = True):
        """n_steps x step(action_sample()) with in-kernel actions; returns the last step's outputs."""
        self._out_buffers(write_obs)
        _lib.check(self._lib.mt_rollout_random(self._h, int(n_steps), self._p(self._obs if write_obs else None),
                                               self._p(self._reward), self._p(self._done), self._stream()))
        return (self._obs if write_obs else None, self._reward, self._done)

    # -- host-buffer (end-to-end) step ------------------------------------------
    def pinned(self, name: str, shape, dtype) -> np.ndarray:
        key = (name, tuple(shape), np.dtype(dtype).str)
        if key not in self._pinned:
            pa = self._pinned[key] = PinnedArray(shape, dtype)
            self._pinned_ptr[id(pa.array)] = pa.ptr
        return self._pinned[key].array

    def step_host(self, actions_host: np.ndarray, write_obs: bool = True):
        """Step with HOST buffers: H2D(actions) + kernel + D2H(obs, reward, done), chunked
        and overlapped on internal streams.  `actions_host` should come from
        `pinned(name, (N, J), float32)` (any name: several buffers can be cycled); other
        input is staged through `pinned('actions', ...)`.  The returned arrays are views of
        page-locked buffers that the NEXT step_host overwrites; they stay valid (they own the
        allocation) even after this object is gone."""
        io = self._host_io.get(write_obs)
        if io is None:                                     # buffers and addresses once: a batch of one is pure latency
            act = self.pinned("actions", (self.n, self.j), np.float32)
            obs = self.pinned("obs", (self.n, 3 * self.x), np.float32) if write_obs else None
            rew = self.pinned("reward", (self.n,), np.float32)
            done = self.pinned("done", (self.n,), np.uint8)
            addr = lambda a: self._pinned_ptr[id(a)] if a is not None else None
            io = self._host_io[write_obs] = (act, addr(act), obs, addr(obs), rew, addr(rew), done, addr(done))
        act, act_ptr, obs, obs_ptr, rew, rew_ptr, done, done_ptr = io
        own = self._pinned_ptr.get(id(actions_host))       # already one of this object's page-locked buffers?
        if own is None:
            np.copyto(act, np.asarray(actions_host, dtype=np.float32).reshape(self.n, self.j))
            own = act_ptr
        _lib.check(self._lib.mt_step_host(self._h, own, obs_ptr, rew_ptr, done_ptr))
        return obs, rew, done

    @property
    def host_step_mode(self) -> str:
        """Which implementation `step_host` settled on for this handle (it times both on its first calls)."""
        return {0: "staged", 1: "zero-copy", 2: "deciding"}.get(int(self._lib.mt_host_step_mode(self._h)), "?")

    # -- state exchange ---------------------------------------------------------
    def set_points(self, points, mask=None):
        p = self._as_dev(points, torch.float32, (self.n, self.x, 3))
        _lib.check(self._lib.mt_set_points(self._h, self._p(p), self._p(self._mask(mask)), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    def get_points(self, zero_dead: bool = True) -> torch.Tensor:
        out = self._empty((self.n, self.x, 3), torch.float32)
        _lib.check(self._lib.mt_get_points(self._h, self._p(out), int(zero_dead), self._stream()))
        return out

    def set_state(self, goals=None, alive=None, total_reward=None, ep_len=None, mask=None):
        g = self._as_dev(goals, torch.float32, (self.n, self.j)) if goals is not None else None
        al = self._alive_words(alive) if alive is not None else None
        tr = self._as_dev(total_reward, torch.float32, (self.n,)) if total_reward is not None else None
        el = self._as_dev(ep_len, torch.int32, (self.n,)) if ep_len is not None else None
        _lib.check(self._lib.mt_set_state(self._h, self._p(g), self._p(al), self._p(tr), self._p(el),
                                          self._p(self._mask(mask)), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    def _alive_words(self, alive) -> torch.Tensor:
        """(N, X) bool or (N,) integer bitmask -> (N,) int32 holding the uint32 bit pattern."""
        al = torch.as_tensor(alive)
        if al.dim() == 2:
            w = (1 << torch.arange(self.x, dtype=torch.int64))
            al = (al.to(torch.int64) * w).sum(dim=1)
        al = al.to(torch.int64) & 0xFFFFFFFF
        al = torch.where(al >= 2 ** 31, al - 2 ** 32, al).to(torch.int32)
        return al.to(self.device).contiguous().reshape(self.n)

    def get_state(self) -> dict:
        g = self._empty((self.n, self.j), torch.float32)
        al = self._empty((self.n,), torch.int32)
        tr = self._empty((self.n,), torch.float32)
        el = self._empty((self.n,), torch.int32)
        _lib.check(self._lib.mt_get_state(self._h, self._p(g), self._p(al), self._p(tr), self._p(el), self._stream()))
        return dict(goals=g, alive=al, total_reward=tr, ep_len=el)

    def alive_matrix(self) -> torch.Tensor:
        """(N, X) bool view of the alive bitmask (the reference's `alives`, manytor.py:134)."""
        al = self.get_state()["alive"].to(torch.int64) & 0xFFFFFFFF
        bits = torch.arange(self.x, device=self.device, dtype=torch.int64)
        return ((al[:, None] >> bits[None, :]) & 1).bool()

    def set_objective_stream(self, points):
        """points (E, N, X, 3): the e-th reset of env n takes set e mod E (parity runs)."""
        if points is None:
            _lib.check(self._lib.mt_set_objective_stream(self._h, None, 0))
            self._stream_keepalive = None
            return
        p = self._as_dev(points, torch.float32)
        assert p.dim() == 4 and tuple(p.shape[1:]) == (self.n, self.x, 3), p.shape
        torch.cuda.current_stream(self.device).synchronize()
        _lib.check(self._lib.mt_set_objective_stream(self._h, self._p(p), int(p.shape[0])))
        self._stream_keepalive = p

    def joints_of(self, goals) -> torch.Tensor:
        """joints_coordinates (manytor.py:188-189) of arbitrary poses (M, J) -> (M, J, 3)."""
        g = self._as_dev(goals, torch.float32).reshape(-1, self.j)
        out = self._empty((g.shape[0], self.j, 3), torch.float32)
        _lib.check(self._lib.mt_joints(self._h, self._p(g), self._p(out), g.shape[0], self._stream()))
        return out

    def fetch_env(self, index: int) -> dict:
        """One-env copy-back for rendering / attribute access (manytor.py:131-139)."""
        goals = np.zeros(self.j, dtype=np.float32)
        joints = np.zeros((self.j, 3), dtype=np.float32)
        points = np.zeros((self.x, 3), dtype=np.float32)
        alive = C.c_uint32()
        total = C.c_float()
        _lib.check(self._lib.mt_fetch_env(self._h, int(index), goals.ctypes.data, joints.ctypes.data,
                                          points.ctypes.data, C.byref(alive), C.byref(total)))
        alives = np.array([(alive.value >> p) & 1 for p in range(self.x)], dtype=bool)
        return dict(goals=goals, joints_coordinates=joints, points=points, alives=alives,
                    total_reward=float(total.value))

    def set_seed(self, seed: int):
        """Re-seed the on-device action / objective streams (gym-style reset(seed=...)): new Philox key,
        per-env episode counters and the step index back to zero, so the same seed reproduces the same
        objectives and actions on the same handle."""
        _lib.check(self._lib.mt_set_seed(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF))
        self.cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF

    @property
    def step_index(self) -> int:
        """Steps executed so far; keys the action stream.  Lives on the device (advanced by the step
        kernel, so CUDA-graph replays count); reading it synchronises."""
        v = C.c_uint64()
        _lib.check(self._lib.mt_get_step_index(self._h, C.byref(v)))
        return int(v.value)

    @step_index.setter
    def step_index(self, value: int):
        _lib.check(self._lib.mt_set_step_index(self._h, int(value) & 0xFFFFFFFFFFFFFFFF))

    # -- statistics ---------------------------------------------------------------
    def stats_tensor(self) -> torch.Tensor:
        """MT_STATS_WORDS int64 on the device (sum-reducible across shards)."""
        out = self._empty((_lib.MT_STATS_WORDS,), torch.int64)
        _lib.check(self._lib.mt_stats_device(self._h, self._p(out), self._stream()))
        return out

    def stats(self) -> dict:
        s = MtStats()
        _lib.check(self._lib.mt_stats_host(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k in STATS_FIELDS}

    def clear_stats(self):
        _lib.check(self._lib.mt_stats_clear(self._h, self._stream()))

    # -- introspection --------------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return int(self._lib.mt_launch_count(self._h))

    def bytes_per_env_step(self, actions_from_hbm: bool = True, obs_written: bool = True) -> int:
        return int(self._lib.mt_bytes_per_env_step(self._h, int(actions_from_hbm), int(obs_written)))

    def set_timing(self, enabled: bool):
        _lib.check(self._lib.mt_set_timing(self._h, int(enabled)))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        _lib.check(self._lib.mt_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)
