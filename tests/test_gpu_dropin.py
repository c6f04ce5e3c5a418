"""The reference-shaped module (manytor_b200.manytor) used the way the reference's own
driver scripts use theirs (test_single.py / test_multi.py), plus API-level behaviour."""
import numpy as np
import pytest

from oracle import OracleEnvs, fk as oracle_fk, dh as oracle_dh, r_theta as oracle_r_theta

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tor():
    import torch
    assert torch.cuda.is_available()
    import manytor_b200.manytor as tor
    return tor


def test_module_level_functions(tor, golden):
    k = golden("known_answers")
    np.testing.assert_allclose(tor.fk(4, [30, 45, 60, 90]), k["fk4_30_45_60_90"], atol=2e-5)   # manytor.py:35-53
    np.testing.assert_allclose(tor.fk(3, [30, 45, 60, 90]), k["fk3_30_45_60_90"], atol=2e-5)
    np.testing.assert_allclose(tor.fk(2, [30, 45, 60, 90]), k["fk2_30_45_60_90"], atol=2e-5)
    g = np.random.RandomState(0).uniform(-180, 180, size=(64, 4))
    np.testing.assert_allclose(tor.fk(4, g), oracle_fk(4, g), atol=3e-5)
    np.testing.assert_allclose(tor.dh(27.0, np.pi / 2, 0.5, 0.3), k["dh_sample"], atol=1e-6)      # manytor.py:25-32
    np.testing.assert_allclose(tor.r_theta([1.0, 2.0, 3.0], [4.0, -1.0, 0.5]), k["r_theta_sample"], atol=1e-4)
    assert tor.r_theta([1.0, 1.0, 1.0], [1.0, 1.0, 1.0]) == (0.0, 0.0)                            # atan2(0, 0) = 0
    assert (tor.HOST, tor.PORT) == ("localhost", 5001)


def test_environment_like_test_single(tor):
    """test_single.py:9-32 with the import swapped, checked against the oracle on the same objectives."""
    env = tor.Environment(10, seed=3)
    obs = env.reset(returnable=True)
    assert obs.shape == (30,) and obs.dtype == np.float64
    ora = OracleEnvs(1, 10)
    ora.reset(points=env.points[None])
    np.testing.assert_allclose(env.joints_coordinates, ora.joints[0], atol=1e-4)
    np.testing.assert_allclose(obs, ora.get_observations()[0], atol=2e-3)
    total = 0
    for _ in range(60):
        action = env.action_sample()
        assert len(action) == 4 and all(isinstance(a, int) and -180 <= a < 180 for a in action)
        obs2, reward, done = env.step(action)
        r = ora.step(np.array([action], dtype=np.float64))
        assert isinstance(reward, int) and isinstance(done, bool) and obs2.shape == (30,)
        assert reward == int(r.reward[0]) and done == bool(r.done[0])
        np.testing.assert_allclose(obs2[0::3], r.obs[0, 0::3], atol=2e-3)
        total += reward
        if done:
            break
    assert env.total_reward == total == ora.total_reward[0]
    np.testing.assert_array_equal(env.alives, ora.alive[0])
    np.testing.assert_allclose(env.goals, ora.goals[0])
    env.reset()
    assert env.total_reward == 0.0 and env.alives.all() and np.all(env.goals == 0)


def test_step_before_reset_raises(tor):
    import manytor_b200
    env = tor.Environment(10)
    with pytest.raises(manytor_b200.MantorLibraryError):      # the reference raises IndexError (manytor.py:143)
        env.step([0, 0, 0, 0])


def test_multienv_like_test_multi(tor):
    """test_multi.py:11-34: lists in, lists out; `done == True` on the list is False (quirk C8)."""
    me = tor.Multienv(env_shape=(3, 2), obj_number=7, seed=5)
    obs = me.reset(returnable=True)
    assert isinstance(obs, list) and len(obs) == 6 and obs[0].shape == (21,)
    assert (me.env_number, me.obj_number, me.rendering) == (6, 7, False)
    for _ in range(20):
        action = me.action_sample()
        assert len(action) == 6 and len(action[0]) == 4
        obs2, reward, done = me.step(action)
        assert isinstance(obs2, list) and isinstance(reward, list) and isinstance(done, list)
        assert all(r in (-1, 0, 1) for r in reward) and all(isinstance(d, bool) for d in done)
        if done == True:  # noqa: E712  (what test_multi.py:22 does)
            break
    totals = [me.environment[i].total_reward for i in range(me.env_number)]      # test_multi.py:32
    assert len(totals) == 6 and all(float(t) == int(t) for t in totals)
    me.reset()
    assert all(me.environment[i].total_reward == 0.0 for i in range(6))


def test_multienv_arrays_and_auto_reset(tor):
    n = 64 * 64
    me = tor.Multienv(env_shape=(64, 64), obj_number=10, as_lists=False, auto_reset=True, horizon=25, seed=2)
    me.reset()
    ended = 0
    for t in range(60):
        a = me.action_sample()
        assert a.shape == (n, 4)
        obs2, reward, done = me.step(a)
        assert obs2.shape == (n, 30) and reward.shape == (n,) and done.shape == (n,)
        ended += int((done != 0).sum())
    st = me.batched.stats()
    assert st["episodes"] == ended and st["env_steps"] == 60 * n
    assert ended >= 2 * n                                   # horizon 25 over 60 steps
    assert st["length_sum"] <= 25 * st["episodes"]


def test_obs_after_reset_option(tor):
    """auto_reset with obs_after_reset=1 returns the first observation of the new episode."""
    from manytor_b200 import BatchedEnvs
    n, x = 2048, 10
    env = BatchedEnvs(n, x, device=0, auto_reset=True, horizon=3, obs_after_reset=True, seed=4)
    env.reset()
    for t in range(3):
        obs, rew, done = env.step(env.sample_actions())
    assert bool((done.cpu().numpy() != 0).all())            # horizon 3: everyone just ended and was reset
    # (same objectives, zero pose; the in-kernel path uses the host-computed zero-pose anchor, so ulps differ)
    np.testing.assert_allclose(obs.cpu().numpy(), env.observe().cpu().numpy(), atol=1e-3)
    st = env.get_state()
    assert np.all(st["goals"].cpu().numpy() == 0) and np.all(st["ep_len"].cpu().numpy() == 0)


def test_million_env_properties(tor):
    """BASELINE config 3 size (2^20 envs): determinism, value ranges, conservation laws."""
    import torch
    from manytor_b200 import BatchedEnvs
    n, x = 1 << 20, 10

    def run(seed):
        env = BatchedEnvs(n, x, device=0, auto_reset=True, horizon=40, seed=seed)
        env.reset()
        pos = neg = 0
        chk = torch.zeros((), dtype=torch.float64, device="cuda")
        for t in range(50):
            obs, rew, done = env.rollout_random(1)
            pos += int((rew == 1).sum())
            neg += int((rew == -1).sum())
            chk += obs.double().sum() + rew.double().sum() * 3 + done.double().sum() * 7
            assert bool(((rew == -1) | (rew == 0) | (rew == 1)).all())
            assert bool((done <= 2).all()) and bool((obs >= 0).all()) and bool(torch.isfinite(obs).all())
            assert bool((obs.view(n, x, 3)[:, :, 1:] <= 90.0001).all())         # both bearings lie in [0, 90]
        return env, pos, neg, float(chk)

    a, pos, neg, chk_a = run(7)
    s = a.stats()
    assert s["env_steps"] == 50 * n
    assert s["episodes"] >= n                                   # horizon 40 < 50 steps
    # every reward ever paid is either in a finished episode or still live
    assert s["reward_sum"] + s["live_reward_sum"] == pos - neg
    assert s["ground_steps"] == neg and s["catches"] >= s["terminated"] * x   # reward -1 <=> ground contact
    assert 0.70 < neg / (50 * n) < 0.88                          # ground-hit rate 0.794
    b, _, _, chk_b = run(7)
    assert chk_a == chk_b and a.stats() == b.stats()             # bit-deterministic
    c, _, _, chk_c = run(8)
    assert chk_c != chk_a


def test_vector_env_adapter_and_dlpack(tor):
    """Gymnasium-style 5-tuple API on device tensors; observations are handed over without a copy."""
    import torch
    from manytor_b200 import ManyTorVectorEnv
    n = 4096
    env = ManyTorVectorEnv(n, 10, max_episode_steps=20, seed=11)
    obs, info = env.reset(seed=123)
    assert obs.shape == (n, 30) and obs.is_cuda and info == {}
    first = obs.clone()
    ended = torch.zeros(n, dtype=torch.bool, device=obs.device)
    for t in range(20):
        obs, reward, terminated, truncated, info = env.step(env.sample_actions())
        assert reward.shape == (n,) and terminated.dtype == torch.bool and truncated.dtype == torch.bool
        assert not bool((terminated & truncated).any())
        ended |= terminated | truncated
    assert bool(truncated[~terminated].all()) and bool(ended.all())          # step 20 = max_episode_steps
    via = torch.utils.dlpack.from_dlpack(obs.__dlpack__())
    assert via.data_ptr() == obs.data_ptr()                                   # zero-copy hand-off
    # same seed -> same first observations and same trajectory of rewards
    env2 = ManyTorVectorEnv(n, 10, max_episode_steps=20, seed=999)
    obs2, _ = env2.reset(seed=123)
    assert torch.equal(obs2, first)
    assert env.episode_statistics()["episodes"] >= n


def test_plain_c_host_runs(tor, tmp_path):
    """The C-ABI used from plain C (examples/c_host.c): test_multi.py's loop with host buffers."""
    import subprocess
    from test_abi import _build_c_host
    exe = _build_c_host(tmp_path)
    res = subprocess.run([exe, "5000", "60"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr + res.stdout
    assert "envs 5000 steps 60" in res.stdout and "300000 env-steps" in res.stdout


@pytest.mark.parametrize("n,x", [(1, 10), (31, 10), (33, 7), (95, 20), (1000, 10), (4097, 10)])
def test_outputs_stay_inside_their_buffers(tor, n, x):
    """compute-sanitizer is closed on this pool, so overruns of the caller's output buffers (the tail
    tile is the risky one: its observations may not leave by a whole-tile bulk store) are caught with
    guard bands: every output lives inside a larger sentinel-filled allocation."""
    import ctypes as C
    import torch
    from manytor_b200 import BatchedEnvs, _lib
    env = BatchedEnvs(n, x, device=0, auto_reset=True, horizon=4, obs_after_reset=True, seed=n)
    env.reset()
    lib = _lib.load()
    G = 4096                                                  # guard elements on each side (16-byte multiples)
    SENT = -12345.0

    def guarded(count, dtype, fill):
        t = torch.full((count + 2 * G,), fill, dtype=dtype, device="cuda")
        return t, t[G:G + count]

    obs_all, obs = guarded(n * 3 * x, torch.float32, SENT)
    rew_all, rew = guarded(n, torch.float32, SENT)
    done_all, done = guarded(n, torch.uint8, 77)
    jn_all, jn = guarded(n * 4 * 3, torch.float32, SENT)
    act_all, act = guarded(n * 4, torch.float32, 0.0)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    for it in range(6):
        _lib.check(lib.mt_sample_actions(env._h, p(act), stream))
        _lib.check(lib.mt_step(env._h, p(act), p(obs), p(rew), p(done), p(jn), stream))
        _lib.check(lib.mt_rollout_random(env._h, 1, p(obs), p(rew), p(done), stream))
        _lib.check(lib.mt_observe(env._h, p(obs), stream))
    torch.cuda.synchronize()
    for whole, fill in ((obs_all, SENT), (rew_all, SENT), (jn_all, SENT), (act_all, 0.0)):
        assert bool((whole[:G] == fill).all()) and bool((whole[-G:] == fill).all())
    assert bool((done_all[:G] == 77).all()) and bool((done_all[-G:] == 77).all())
    assert bool((obs != SENT).all()) and bool((rew != SENT).all()) and bool((done <= 2).all()) and bool((jn != SENT).all())


def test_remaining_entry_points(tor):
    """step_host without observations, stats clear, config read-back, seeds."""
    import ctypes as C
    import torch
    from manytor_b200 import BatchedEnvs, _lib
    n = 3000
    a = BatchedEnvs(n, 10, device=0, seed=1, auto_reset=True, horizon=5)
    b = BatchedEnvs(n, 10, device=0, seed=1, auto_reset=True, horizon=5)
    a.reset(); b.reset()
    rng = np.random.RandomState(0)
    for _ in range(7):
        act = rng.randint(-180, 180, size=(n, 4)).astype(np.float32)
        oa, ra, da = a.step_host(act, write_obs=False)
        ob, rb, db = b.step_host(act)
        assert oa is None
        np.testing.assert_array_equal(ra, rb)
        np.testing.assert_array_equal(da, db)
    assert a.stats() == b.stats() and a.stats()["episodes"] >= n
    a.clear_stats()
    s = a.stats()
    assert s["episodes"] == 0 and s["env_steps"] == 0 and s["reward_sum"] == 0
    cfg = _lib.MtConfig()
    _lib.check(_lib.load().mt_get_config(a._h, C.byref(cfg)))
    assert (cfg.n_envs, cfg.n_obj, cfg.horizon, cfg.auto_reset, cfg.seed) == (n, 10, 5, 1, 1)
    # re-seeding changes the action stream from the next launch on
    x0 = a.sample_actions().clone()
    a.set_seed(2)
    x1 = a.sample_actions()
    assert not torch.equal(x0, x1)
