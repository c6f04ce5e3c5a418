"""Known-answer vectors for the oracle's Philox4x32-10 (Random123 kat_vectors)."""
import numpy as np

from oracle.philox import device_actions, philox4x32_10


def _one(c, k):
    r = philox4x32_10(*[np.array([v], dtype=np.uint32) for v in c], *[np.array([v], dtype=np.uint32) for v in k])
    return [int(v[0]) for v in r]


def test_random123_known_answers():
    assert _one([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _one([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _one([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_action_map_range_and_uniformity():
    a = device_actions(42, np.arange(200000), 3, 4)
    assert a.min() == -180 and a.max() == 179          # manytor.py:216: integers on [-180, 180)
    counts = np.bincount((a + 180).ravel(), minlength=360)
    assert counts.min() > 0.9 * a.size / 360 and counts.max() < 1.1 * a.size / 360
    assert not np.array_equal(a, device_actions(42, np.arange(200000), 4, 4))
    np.testing.assert_array_equal(a[1000:2000], device_actions(42, np.arange(1000, 2000), 3, 4))  # shard-invariant
