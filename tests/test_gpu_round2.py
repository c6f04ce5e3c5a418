"""Round-2 behaviour of the C ABI and the drop-in module (-m gpu): device-side step counter under
CUDA-graph replay, re-seeding, the first-reset rule, truncation in the drop-in's `done`, per-env
methods and `trajectory` of `Multienv.environment[i]`, the Gymnasium adapter against the oracle,
the zero-copy host step, the NCCL statistics all-reduce of the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import OracleEnvs, REFERENCE_ARM
from oracle.manytor_oracle import _route, fk_frames
from oracle.philox import device_actions

from parity import Report, alive_bits_to_matrix, compare_step, half_ball_points

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mt():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import manytor_b200
    manytor_b200.load_library()
    return manytor_b200


# --------------------------------------------------------------------------
# the step index and the statistics live on the device
# --------------------------------------------------------------------------
@pytest.mark.parametrize("n", [200_003, 5_000])
def test_replayed_random_rollout_graph_equals_eager(mt, n):
    """A captured mt_rollout_random replayed R times must be the same 2R*k steps as the eager calls: the
    step index that keys the action stream is read from device memory and advanced by the kernel, so a
    replay draws FRESH actions and its env-steps are counted (both were host-side in round 1)."""
    import torch
    x, k, reps = 10, 6, 3
    kw = dict(device=0, seed=5, auto_reset=True, horizon=9)
    eager = mt.BatchedEnvs(n, x, **kw)
    eager.reset()
    outs = []
    for _ in range(reps * k):
        o, r, d = eager.rollout_random(1)
        outs.append((o.clone(), r.clone(), d.clone()))
    cap = mt.BatchedEnvs(n, x, **kw)
    cap.reset()
    cap._out_buffers(True, False)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = cap.rollout_random(k)
    # capture launched nothing: the counters have not moved
    assert cap.step_index == 0 and cap.stats()["env_steps"] == 0
    for rep in range(reps):
        g.replay()
        torch.cuda.synchronize()
        for u, v in zip(out, outs[(rep + 1) * k - 1]):
            assert torch.equal(u, v), f"replay {rep} differs from the eager run"
    assert cap.step_index == eager.step_index == reps * k
    assert cap.stats() == eager.stats() and cap.stats()["env_steps"] == n * reps * k
    se, sc = eager.get_state(), cap.get_state()
    for key in se:
        assert torch.equal(se[key], sc[key]), key
    # and the stream of actions is the documented one: Philox keyed by (seed, env id, step index)
    a = cap.sample_actions().cpu().numpy().astype(np.int64)
    np.testing.assert_array_equal(a[:64], device_actions(5, np.arange(64), reps * k, 4))


def test_step_index_checkpoint(mt):
    import torch
    n = 3000
    a = mt.BatchedEnvs(n, 10, device=0, seed=3)
    a.reset()
    a.rollout_random(4)
    assert a.step_index == 4
    snap = a.sample_actions().clone()
    a.rollout_random(3)
    a.step_index = 4                                          # resume: the same actions again
    assert torch.equal(a.sample_actions(), snap)
    a.step(snap)                                              # mt_step advances it too
    assert a.step_index == 5
    a.step_host(snap.cpu().numpy())                           # and the chunked host step, once per call
    assert a.step_index == 6 and a.stats()["env_steps"] == n * 9


def test_reseeding_reproduces_objectives_and_actions(mt):
    """reset(seed=s) twice on ONE handle gives the same objectives and actions (ADVICE: only the Philox
    key used to change; the episode counters and the step index kept running)."""
    import torch
    env = mt.BatchedEnvs(2000, 10, device=0, seed=1, auto_reset=True, horizon=4)
    env.set_seed(42)
    env.reset()
    p0 = env.get_points(False).clone()
    tr0 = [tuple(t.clone() for t in env.rollout_random(1)) for _ in range(9)]
    env.set_seed(42)
    env.reset()
    assert torch.equal(env.get_points(False), p0)
    for want in tr0:
        got = env.rollout_random(1)
        assert all(torch.equal(u, v) for u, v in zip(got, want))
    v = mt.ManyTorVectorEnv(512, 10, max_episode_steps=5, seed=0)
    o1, _ = v.reset(seed=7)
    o1 = o1.clone()
    a1 = v.sample_actions().clone()
    for _ in range(7):
        v.step(v.sample_actions())
    o2, _ = v.reset(seed=7)
    assert torch.equal(o1, o2) and torch.equal(a1, v.sample_actions())


def test_first_reset_must_cover_every_env(mt):
    env = mt.BatchedEnvs(100, 10, device=0)
    m = np.ones(100, dtype=bool)
    m[37] = False
    with pytest.raises(mt.MantorLibraryError, match="first mt_reset must cover every env"):
        env.reset(mask=m)
    with pytest.raises(mt.MantorLibraryError, match="before mt_reset"):
        env.step(np.zeros((100, 4), dtype=np.float32))
    env.reset(mask=np.ones(100, dtype=bool))                  # an all-ones mask is a full reset
    env.reset(mask=m)                                         # partial resets are fine afterwards
    env.step(np.zeros((100, 4), dtype=np.float32))


# --------------------------------------------------------------------------
# drop-in module
# --------------------------------------------------------------------------
def test_dropin_done_reports_truncation(mt):
    import manytor_b200.manytor as tor
    env = tor.Environment(10, seed=1, horizon=3, auto_reset=True)
    env.reset()
    flags = [env.step(env.action_sample())[2] for _ in range(6)]
    assert flags == [False, False, True, False, False, True]          # the horizon ends episodes visibly
    me = tor.Multienv((2, 2), 5, seed=1, horizon=2, auto_reset=True)
    me.reset()
    d = [me.step(me.action_sample())[2] for _ in range(4)]
    assert d[1] == [True] * 4 and d[3] == [True] * 4 and d[0] == [False] * 4
    # no horizon: the reference's meaning, all objectives collected
    me = tor.Multienv((2, 2), 5, seed=1)
    me.reset()
    assert me.step(me.action_sample())[2] == [False] * 4


def _trajectory_oracle(pairs):
    rows = [np.array([[0.0, 0.0, 51.3]])]
    for before, action in pairs:
        route = _route(np.asarray(before, dtype=np.float64)[None], np.asarray(action, dtype=np.float64)[None], 25)
        rows.append(np.stack([fk_frames(route[p], REFERENCE_ARM)[0, 4] for p in range(25)]))   # manytor.py:188-190
    return np.vstack(rows)


def test_trajectory_matches_reference_semantics(mt):
    """manytor.py:135,190,223: [0,0,51.3], then the terminal at each of the 25 sub-poses of every step."""
    import manytor_b200.manytor as tor
    env = tor.Environment(10, seed=2)
    env.reset()
    np.testing.assert_array_equal(env.trajectory, [0.0, 0.0, 51.3])
    pairs, pose = [], np.zeros(4)
    for t in range(5):
        a = env.action_sample()
        env.step(a)
        pairs.append((pose, np.array(a, dtype=np.float64)))
        pose = np.array(a, dtype=np.float64)
        if t == 2:
            assert env.trajectory.shape == (1 + 25 * 3, 3)                   # read in the middle, then extended
    want = _trajectory_oracle(pairs)
    assert env.trajectory.shape == want.shape == (126, 3)
    np.testing.assert_allclose(env.trajectory, want, atol=5.6e-4)
    env.reset()
    np.testing.assert_array_equal(env.trajectory, [0.0, 0.0, 51.3])


def test_envview_is_a_real_environment(mt):
    """manytor.py:82,88-89: `Multienv.environment[i]` holds Environments with step/reset/render and a
    trajectory.  A per-env step must move only that env, exactly like the oracle's single env."""
    import manytor_b200.manytor as tor
    me = tor.Multienv((2, 3), 6, seed=9)
    me.reset()
    pts = np.stack([me.environment[i].points for i in range(6)])
    ora = OracleEnvs(6, 6)
    ora.reset(points=pts)
    acts = me.action_sample()
    me.step(acts)
    ora.step(np.array(acts, dtype=np.float64))
    e2 = me.environment[2]
    before = [me.environment[i].goals for i in range(6)]
    a = [10, -70, 33, 120]
    solo = OracleEnvs(1, 6)
    solo.reset(points=ora.points[2:3])
    solo.goals, solo.alive, solo.total_reward = ora.goals[2:3].copy(), ora.alive[2:3].copy(), ora.total_reward[2:3].copy()
    r = solo.step(np.array([a], dtype=np.float64))
    obs, rew, done = e2.step(a)
    assert rew == int(r.reward[0]) and done == bool(r.done[0])
    np.testing.assert_allclose(obs[0::3], r.obs[0, 0::3], atol=2e-3)
    np.testing.assert_allclose(e2.goals, a)
    np.testing.assert_array_equal(e2.alives, r.alive[0])
    assert e2.total_reward == float(solo.total_reward[0])
    for i in (0, 1, 3, 4, 5):                                               # nobody else moved
        np.testing.assert_array_equal(me.environment[i].goals, before[i])
    want = _trajectory_oracle([(np.zeros(4), np.array(acts[2], dtype=np.float64)),
                               (np.array(acts[2], dtype=np.float64), np.array(a, dtype=np.float64))])
    np.testing.assert_allclose(e2.trajectory, want, atol=5.6e-4)
    np.testing.assert_allclose(me.environment[0].trajectory,
                               _trajectory_oracle([(np.zeros(4), np.array(acts[0], dtype=np.float64))]), atol=5.6e-4)
    assert len(e2.action_sample()) == 4 and e2.get_observations().shape == (18,) and e2.is_done() is False
    o = e2.reset(returnable=True)
    assert o.shape == (18,) and np.all(e2.goals == 0) and e2.total_reward == 0.0 and e2.alives.all()
    np.testing.assert_array_equal(e2.trajectory, [0.0, 0.0, 51.3])
    np.testing.assert_array_equal(me.environment[1].goals, before[1])       # the reset touched env 2 only
    e2.render()
    assert 2 in me.render_envs
    e2.render(stop_render=True)
    assert 2 not in me.render_envs


def test_action_sample_is_the_reference_stream(mt):
    """manytor.py:215-217, 111-113: action_sample() draws np.random.randint(-180, 180) per joint from the
    process-global stream, env after env -- a caller that seeds np.random gets the reference's actions."""
    import manytor_b200.manytor as tor
    from oracle import action_sample_reference_stream
    want = lambda k: [[int(v) for v in action_sample_reference_stream()] for _ in range(k)]
    env = tor.Environment(10)
    env.reset()
    np.random.seed(11)
    got = [env.action_sample() for _ in range(5)]
    np.random.seed(11)
    assert got == want(5) and all(type(v) is int for v in got[0])
    me = tor.Multienv((2, 3), 5)
    me.reset()
    np.random.seed(12)
    got = me.action_sample() + [me.environment[4].action_sample()]
    np.random.seed(12)
    assert got == want(7)
    arr = tor.Multienv((2, 3), 5, as_lists=False)                          # array mode: the device sampler
    arr.reset()
    a = arr.action_sample()
    assert a.shape == (6, 4) and a.dtype == np.int64 and a.min() >= -180 and a.max() < 180


def test_pinned_views_outlive_their_env(mt):
    """ADVICE: step_host() returns views of page-locked buffers; they must stay valid after the env that
    allocated them is gone (the views keep the allocation alive)."""
    import gc
    env = mt.BatchedEnvs(5000, 10, device=0, seed=3)
    env.reset()
    act = np.random.RandomState(0).randint(-180, 180, size=(5000, 4)).astype(np.float32)
    obs, rew, done = env.step_host(act)
    keep = (obs.copy(), rew.copy(), done.copy())
    del env
    gc.collect()
    junk = [mt.BatchedEnvs(4096, 10, device=0) for _ in range(3)]          # churn the allocator
    for j in junk:
        j.reset()
        j.step_host(act[:4096])
    np.testing.assert_array_equal(obs, keep[0])
    np.testing.assert_array_equal(rew, keep[1])
    np.testing.assert_array_equal(done, keep[2])


# --------------------------------------------------------------------------
# Gymnasium adapter against the oracle
# --------------------------------------------------------------------------
@pytest.mark.parametrize("x", [1, 4])
def test_vector_env_values_against_oracle(mt, x):
    """The 5-tuple of ManyTorVectorEnv.step compared VALUE by value with the oracle, including the rows
    of envs that ended: same-step auto-reset returns the first observation of the next episode
    (obs_after_reset), whose objectives come from the uploaded stream here so the oracle can follow."""
    n, horizon, steps, sets = 1024, 12, 60, 8         # x = 1: episodes also end by collecting everything
    rs = np.random.RandomState(4 + x)
    stream = np.float32(half_ball_points(rs, (sets, n, x))).astype(np.float64)
    env = mt.ManyTorVectorEnv(n, x, max_episode_steps=horizon, seed=1)
    env.envs.set_objective_stream(stream)
    obs, info = env.reset()
    ora = OracleEnvs(n, x)
    ora.reset(points=stream[0])
    np.testing.assert_allclose(obs.cpu().numpy()[:, 0::3], ora.get_observations()[:, 0::3], atol=2e-3)
    episode = np.ones(n, dtype=np.int64)
    rep = Report()
    n_term = n_trunc = 0
    for t in range(steps):
        act = rs.randint(-180, 180, size=(n, 4)).astype(np.float64)
        alive_before, points_before = ora.alive.copy(), ora.points.copy()
        import torch
        o, rew, term, trunc, info = env.step(torch.as_tensor(act, dtype=torch.float32, device="cuda"))
        o, rew, term, trunc = o.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
        r = ora.step(act)
        want_trunc = (ora.ep_len >= horizon) & ~r.done
        ended = r.done | want_trunc
        near = (r.ground_margin < 1e-3) | (np.where(alive_before, np.abs(r.catch_margin), np.inf).min(axis=1) < 1e-3)
        ok = (rew.astype(np.int64) == r.reward) & (term == r.done) & (trunc == want_trunc)
        assert (ok | near).all(), "flags differ away from a threshold"
        if not ok.all():
            pytest.skip("a near-threshold flip forked the episode structure; covered by the lock-step tests")
        # envs that go on: obs2 of this step (manytor.py:204)
        go = ~ended
        st = env.envs.get_state()
        dev_alive = np.where(ended[:, None], r.alive, alive_bits_to_matrix(st["alive"].cpu().numpy(), x))
        o_cmp = np.where(go[:, None], o, np.float32(r.obs))                  # compare_step sees oracle rows for ended envs
        compare_step(rep, REFERENCE_ARM, ora, r, points_before, alive_before, o_cmp, rew, term.astype(np.uint8), dev_alive)
        if ended.any():
            fresh = stream[episode % sets, np.arange(n)]
            ora.reset(mask=ended, points=fresh)
            episode[ended] += 1
            first = ora.get_observations()                                   # manytor.py:251-253 of the new episode
            np.testing.assert_allclose(o[ended][:, 0::3], first[ended][:, 0::3], atol=2e-3)
            d = np.abs(o[ended] - first[ended]).reshape(-1, x, 3)
            assert d[..., 1:].max() < 0.05, "bearings of the first observation after reset"
            n_term += int(r.done.sum()); n_trunc += int(want_trunc.sum())
    assert rep.ok(), rep.notes[:5]
    assert n_term + n_trunc >= n * (steps // horizon) and (n_term > 0 or x > 1)
    s = env.episode_statistics()
    assert s["episodes"] == n_term + n_trunc and s["terminated"] == n_term


# --------------------------------------------------------------------------
# host-buffer step: zero-copy variant; NCCL all-reduce of the C ABI
# --------------------------------------------------------------------------
def test_zero_copy_host_step_equals_staged(mt):
    n, x = 70_001, 10
    act = np.random.RandomState(1).randint(-180, 180, size=(6, n, 4)).astype(np.float32)
    res = {}
    for mode in ("0", "1"):
        os.environ["MT_HOST_ZEROCOPY"] = mode
        try:
            env = mt.BatchedEnvs(n, x, device=0, seed=3, auto_reset=True, horizon=4)
            env.reset()
            acc = []
            for t in range(6):
                a = env.pinned("actions", (n, 4), np.float32)
                a[:] = act[t]
                o, r, d = env.step_host(a)
                acc.append((o.copy(), r.copy(), d.copy()))
            res[mode] = (acc, env.stats(), env.step_index)
        finally:
            os.environ.pop("MT_HOST_ZEROCOPY", None)
    for (o0, r0, d0), (o1, r1, d1) in zip(res["0"][0], res["1"][0]):
        np.testing.assert_array_equal(o0, o1)
        np.testing.assert_array_equal(r0, r1)
        np.testing.assert_array_equal(d0, d1)
    assert res["0"][1] == res["1"][1] and res["0"][2] == res["1"][2] == 6


def test_host_step_picks_a_variant_by_itself(mt):
    """Without MT_HOST_ZEROCOPY the handle tries both implementations on its first six calls and keeps one; every
    call -- whichever variant served it -- must equal the device step."""
    n = 40_000
    a = mt.BatchedEnvs(n, 10, device=0, seed=3, auto_reset=True, horizon=3)
    b = mt.BatchedEnvs(n, 10, device=0, seed=3, auto_reset=True, horizon=3)
    a.reset(); b.reset()
    assert a.host_step_mode == "deciding"
    rng = np.random.RandomState(1)
    for t in range(9):
        act = rng.randint(-180, 180, size=(n, 4)).astype(np.float32)
        oa, ra, da = a.step_host(act)
        ob, rb, db = b.step(act)
        np.testing.assert_array_equal(oa, ob.cpu().numpy())
        np.testing.assert_array_equal(ra, rb.cpu().numpy())
        np.testing.assert_array_equal(da, db.cpu().numpy())
    assert a.host_step_mode in ("staged", "zero-copy")
    assert a.stats() == b.stats() and a.step_index == b.step_index == 9


def test_stats_allreduce_c_abi(mt):
    """mt_stats_allreduce: n = 1 is the plain host read; n >= 2 (one handle per device) sums over NCCL."""
    import torch
    from manytor_b200 import _lib
    lib = _lib.load()
    ndev = min(torch.cuda.device_count(), 4)
    envs = [mt.BatchedEnvs(10_000 + 32 * i, 10, device=i, seed=1, auto_reset=True, horizon=5, env_id_base=100_000 * i)
            for i in range(ndev)]
    for e in envs:
        with torch.cuda.device(e.device):
            e.reset()
            e.rollout_random(12)
    per = [e.stats() for e in envs]
    arr = (C.c_void_p * ndev)(*[e._h.value for e in envs])
    out = _lib.MtStats()
    _lib.check(lib.mt_stats_allreduce(arr, ndev, C.byref(out)))
    for k in _lib.STATS_FIELDS:
        assert getattr(out, k) == sum(p[k] for p in per), k
    assert out.env_steps == sum(12 * e.n for e in envs)
    if ndev >= 2:
        dup = (C.c_void_p * 2)(envs[0]._h.value, envs[0]._h.value)
        assert lib.mt_stats_allreduce(dup, 2, C.byref(out)) == -1           # two handles on one device


def test_multi_gpu_c_host_runs(mt, tmp_path):
    """examples/c_host_multi.c: BASELINE config 4 driven from plain C -- one handle per GPU, no per-step
    exchange, one ncclAllReduce of the statistics at the end (mt_stats_allreduce).  Uses every GPU of the box
    (on a single-GPU box the collective degenerates to the host read, which the program checks the same way)."""
    import subprocess
    import torch
    from test_abi import _build_c_host
    exe = _build_c_host(tmp_path, "c_host_multi")
    g = min(torch.cuda.device_count(), 8)
    res = subprocess.run([exe, str(g), "20000", "30"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr + res.stdout
    assert f"gpus {g} x envs 20000 x steps 30: {g * 20000 * 30} env-steps" in res.stdout


# --------------------------------------------------------------------------
# multi-step rollout in one launch (rollout_kernel) == the same steps as separate launches
# --------------------------------------------------------------------------
@pytest.mark.parametrize("arm_name,n,x,obs,oar", [("ref", 150_003, 10, True, False), ("ref", 40_000, 10, False, False),
                                                  ("ref", 5_000, 7, True, True), ("ref", 33, 32, True, False),
                                                  ("ur5", 90_001, 20, True, True), ("ur5", 3_000, 5, False, False),
                                                  ("ref-jit", 9_000, 7, True, False), ("custom-jit", 70_000, 12, True, True)])
def test_multi_step_rollout_equals_per_step_launches(mt, arm_name, n, x, obs, oar):
    """mt_rollout_random(k) keeps a tile's pose / alive word / total reward in registers and its objectives in shared
    memory for all k steps (one launch); MT_ROLLOUT_PERSISTENT=0 forces k launches of the step kernel.  Outputs of
    the last step, state, objectives, statistics and the step index must be bit-identical -- through in-kernel
    resets (horizon 4 < k), the first-observation-after-reset option, ragged sizes, odd / even / maximum objective
    counts, and both built-in arms."""
    import torch
    jit = arm_name.endswith("jit")            # NVRTC-specialised kernels (another objective count / a user's DH table)
    custom = mt.ArmSpec(dh=((0.0, np.pi / 2, 0.30, 0.0), (0.50, 0.0, 0.0, 0.2), (0.40, 0.3, 0.10, 0.0),
                            (0.0, -np.pi / 2, 0.20, -np.pi / 2), (0.10, 0.0, 0.05, 0.0)),
                        obs_frame=4, ground_frames=(4, 5), catch_frame=5, radius=0.9, catch_tol=0.15)
    arm = mt.UR5_ARM if arm_name == "ur5" else (custom if arm_name == "custom-jit" else mt.REFERENCE_ARM)
    kw = dict(arm=arm, device=0, seed=13, auto_reset=True, horizon=4, obs_after_reset=oar)
    res = []
    for persistent in ("1", "0"):
        os.environ["MT_ROLLOUT_PERSISTENT"] = persistent
        if not jit:
            os.environ["MT_DISABLE_JIT"] = "1"     # built-in kernels with a run-time objective count
        try:
            env = mt.BatchedEnvs(n, x, **kw)
            env.reset()
            l0 = env.launch_count
            outs = []
            for k in (7, 2, 11):
                o, r, d = env.rollout_random(k, write_obs=obs)
                outs.append((o.clone() if obs else None, r.clone(), d.clone()))
            launches = env.launch_count - l0
            res.append((outs, {k2: v.clone() for k2, v in env.get_state().items()}, env.get_points(False).clone(), env.stats(),
                        env.step_index, launches))
        finally:
            os.environ.pop("MT_ROLLOUT_PERSISTENT", None)
            os.environ.pop("MT_DISABLE_JIT", None)
    (oa, sa, pa, ta, ia, la), (ob, sb, pb, tb, ib, lb) = res
    assert la == 3 and lb == 20, (la, lb)                     # one launch per call against one per step
    for (o1, r1, d1), (o2, r2, d2) in zip(oa, ob):
        if obs:
            assert torch.equal(o1, o2)
        assert torch.equal(r1, r2) and torch.equal(d1, d2)
    for key in sa:
        assert torch.equal(sa[key], sb[key]), key
    assert torch.equal(pa, pb) and ta == tb and ia == ib == 20
    assert ta["env_steps"] == 20 * n and ta["episodes"] >= 4 * n


def test_policy_loop_example_feeds_fresh_observations(mt):
    """examples/policy_loop.py (the zero-copy hand-off pattern INTEGRATION.md points to): reset() and step() must
    return the SAME observation storage, so that a policy -- eager or captured in a CUDA graph -- keeps reading
    fresh observations (it used to read the reset observations forever).  The example asserts it."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "examples", "policy_loop.py"), "8192", "30"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:] + res.stdout[-500:]
    assert "graph :" in res.stdout and "episode statistics" in res.stdout


@pytest.mark.parametrize("arm_name,x,lo,hi", [("ref", 10, -180, 180), ("ref", 10, -90, 45), ("ref", 7, -1000, 1000),
                                              ("ur5", 20, -180, 180), ("ur5", 6, -30, 400)])
def test_action_sincos_table_is_a_pure_memoisation(mt, arm_name, x, lo, hi):
    """In-kernel actions are integer degrees, so the kernels look the sin / cos of the joint targets up in a per-block
    table filled by the very evaluation they would otherwise run per env and step (ActionTrig, mt_step.cuh).  With the
    table (default), without it (MT_ACTION_TABLE=0), and through mt_step on the sampled actions (which always
    evaluates), every output must be bit-identical -- also for other action ranges and for a range too wide to tabulate."""
    import torch
    arm = mt.UR5_ARM if arm_name == "ur5" else mt.REFERENCE_ARM
    n = 20_000
    kw = dict(arm=arm, device=0, seed=21, auto_reset=True, horizon=6, action_low=lo, action_high=hi)
    runs = []
    for table in ("1", "0", "step"):
        os.environ["MT_ACTION_TABLE"] = "0" if table == "0" else "1"
        try:
            env = mt.BatchedEnvs(n, x, **kw)
            env.reset()
            outs = []
            for t in range(4):
                if table == "step":
                    a = env.sample_actions()
                    assert int(a.min()) >= lo and int(a.max()) < hi
                    o, r, d = env.step(a)
                else:
                    o, r, d = env.rollout_random(1)
                outs.append((o.clone(), r.clone(), d.clone()))
            if table != "step":
                o, r, d = env.rollout_random(9)                    # the multi-step kernel
                outs.append((o.clone(), r.clone(), d.clone()))
            runs.append((outs, {k: v.clone() for k, v in env.get_state().items()}, env.stats()))
        finally:
            os.environ.pop("MT_ACTION_TABLE", None)
    for other in runs[1:]:
        for (o1, r1, d1), (o2, r2, d2) in zip(runs[0][0], other[0]):
            assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
    for key in runs[0][1]:
        assert torch.equal(runs[0][1][key], runs[1][1][key]), key
    assert runs[0][2] == runs[1][2]


def _peer_stats_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import manytor_b200
    from manytor_b200 import distributed as mtd
    mtd.init_from_env()
    torch.cuda.set_device(rank)
    env = manytor_b200.BatchedEnvs(30_000 + 64 * rank, 10, device=rank, seed=2, auto_reset=True, horizon=5, env_id_base=10 ** 6 * rank)
    env.reset()
    red = mtd.StatsReducer(torch.device("cuda", rank))
    sums = []
    for it in range(5):                                           # several exchanges: both slot parities, epoch bookkeeping
        env.rollout_random(7)
        mine = env.stats_tensor().clone()
        total = red.reduce(env)
        ref = mine.clone()
        dist.all_reduce(ref)                                      # NCCL as the checker
        torch.cuda.synchronize()
        sums.append((total.tolist(), ref.tolist()))
    out[rank] = (red.path, sums)
    dist.barrier()
    dist.destroy_process_group()


def test_peer_memory_stats_allreduce_equals_nccl(mt):
    """mt_stats_allreduce_peers: the library's own kernel over NVLink peer memory (buffers mapped by torch symmetric
    memory) must give every rank the same sum as ncclAllReduce, call after call.  Needs >= 2 GPUs."""
    import socket
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_peer_stats_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        path, sums = out[r]
        assert path.startswith("peer memory kernel"), path
        for total, ref in sums:
            assert total == ref
    assert out[0][1] == out[1][1]
