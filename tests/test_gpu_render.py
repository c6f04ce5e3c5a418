"""The render shim: frames on the reference's UDP/JSON protocol (manytor.py:196-201, plotting.py:27-87)
produced from a one-env copy-back, with the 25 sub-pose joint positions computed on the GPU."""
import json
import socket

import numpy as np
import pytest

from oracle.manytor_oracle import joints_coordinates

pytestmark = pytest.mark.gpu


def test_render_frames_follow_reference_protocol(monkeypatch):
    import torch
    assert torch.cuda.is_available()
    import manytor_b200.manytor as tor

    rx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    rx.bind(("127.0.0.1", 0))
    rx.settimeout(5.0)
    monkeypatch.setattr(tor, "PORT", rx.getsockname()[1])
    monkeypatch.setattr(tor, "HOST", "127.0.0.1")
    monkeypatch.setattr(tor, "plot_vispy", lambda: None)          # no viewer process in the test
    monkeypatch.setattr(tor.time, "sleep", lambda s: None)

    x = 5
    env = tor.Environment(x, seed=9)
    env.reset()
    env.render()
    init = json.loads(rx.recvfrom(65536)[0])
    assert init == [1, x, 3]                                       # manytor.py:271-274
    goals0 = env.goals.copy()
    points0 = env.points.copy()
    action = [30, -45, 60, 10]
    env.step(action)
    route = np.linspace(goals0, np.asarray(action, dtype=np.float64), num=25)      # manytor.py:182
    for p in range(25):
        msg = np.array(json.loads(rx.recvfrom(65536)[0]), dtype=np.float64).reshape(-1, 3)
        assert msg.shape == (1 + 4 + x + 1, 3)                     # index, joints, points, trajectory tail
        assert msg[0, 0] == env.id and np.isnan(msg[0, 1]) and msg[0, 2] == (1 if p == 0 else 0)
        np.testing.assert_allclose(msg[1:5], joints_coordinates(route[p]), atol=5.6e-4)
        np.testing.assert_allclose(msg[5:5 + x], points0, atol=1e-6)
        np.testing.assert_allclose(msg[-1], msg[4])               # trajectory tail = end effector
    env.reset()
    clear = json.loads(rx.recvfrom(65536)[0].decode().replace("NaN", "null"))
    assert clear[2] == 4                                           # manytor.py:246-249
    env.render(stop_render=True)
    stop = json.loads(rx.recvfrom(65536)[0].decode().replace("NaN", "null"))
    assert stop[2] == 2                                            # manytor.py:277-279
    rx.close()
