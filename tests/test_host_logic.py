"""Host-side logic that needs no GPU: sharding arithmetic and the one collective
(statistics all-reduce), exercised with world_size=2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from manytor_b200 import distributed as mtd
from manytor_b200._lib import MT_STATS_WORDS, STATS_FIELDS


def test_shard_range_partitions_exactly():
    for n in (1, 7, 8, 1 << 20, (1 << 23) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [mtd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0
            for a, b in zip(spans, spans[1:]):
                assert a[0] + a[1] == b[0]
            assert spans[-1][0] + spans[-1][1] == n
            sizes = [s[1] for s in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        mtd.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = mtd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    base, count = mtd.shard_range(n_total, rank, world)
    # a fake per-shard statistics vector in the layout of mt_stats (include/manytor_b200.h)
    stats = torch.tensor([count * 10, count // 3, count // 7, -count, count * 5, count * 2, count, base],
                         dtype=torch.int64)
    assert stats.numel() == MT_STATS_WORDS
    mtd.allreduce_stats(stats)

    class FakeEnv:                                   # StatsReducer on a CPU group: the all_reduce fallback
        device = torch.device("cpu")

        def stats_tensor(self):
            return torch.tensor([count, 1, 0, 0, 0, 0, 0, rank], dtype=torch.int64)

    red = mtd.StatsReducer(device=None)
    assert red.path == "all_reduce"
    got = red.reduce(FakeEnv()).tolist()
    assert got[0] == n_total and got[1] == world and got[7] == sum(range(world))
    t = mtd.max_over_ranks(1.0 + rank)
    out[rank] = (stats.tolist(), t, base, count)
    dist.barrier()
    dist.destroy_process_group()


def test_stats_allreduce_world2_gloo():
    world, n_total = 2, 1001
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
    counts = [out[r][3] for r in range(world)]
    bases = [out[r][2] for r in range(world)]
    assert sum(counts) == n_total and bases == [0, counts[0]]
    expect = [sum(c * 10 for c in counts), sum(c // 3 for c in counts), sum(c // 7 for c in counts),
              -n_total, n_total * 5, n_total * 2, n_total, sum(bases)]
    for r in range(world):
        assert out[r][0] == expect            # every rank holds the global sum
        assert out[r][1] == 2.0               # max over ranks of the per-rank time
    assert mtd.stats_dict(torch.tensor(expect))["env_steps"] == expect[0]
    assert list(mtd.stats_dict(torch.tensor(expect))) == list(STATS_FIELDS)


def test_pinned_views_keep_their_allocation_alive():
    """arrays handed out over page-locked memory must own it -- the block is released when the LAST
    view goes, not when the wrapper (or the BatchedEnvs that made it) is collected.  Checked here with a fake
    allocator; the GPU suite repeats it on real pinned memory."""
    import ctypes as C
    import gc
    from manytor_b200.core import PinnedArray
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    freed = []
    pa = PinnedArray((4, 3), np.float32, lambda n: libc.malloc(n), lambda p: freed.append(p.value))
    view = pa.array[1:3]
    ptr = pa.ptr
    del pa
    gc.collect()
    assert freed == []
    view[:] = 1.0
    assert float(view.sum()) == 6.0
    del view
    gc.collect()
    assert freed == [ptr]
    with pytest.raises(ValueError):
        PinnedArray((2,), np.float32, alloc=lambda n: libc.malloc(n))


def test_host_action_sample_is_the_reference_stream():
    """manytor.py:215-217 / 111-113: the drop-in's action_sample() is J x np.random.randint(low, high) per env from
    the process-global stream, env after env; one vectorised call must consume that stream like the scalar calls."""
    import manytor_b200
    from manytor_b200.core import make_config
    from manytor_b200.manytor import _host_action_sample
    from oracle import action_sample_reference_stream
    cfg = make_config(6, 5, manytor_b200.REFERENCE_ARM, 0)
    np.random.seed(3)
    got = _host_action_sample(cfg, 6) + _host_action_sample(cfg, 1)
    np.random.seed(3)
    want = [[int(v) for v in action_sample_reference_stream()] for _ in range(7)]
    assert got == want and all(type(v) is int and -180 <= v < 180 for row in got for v in row)
    ur5 = make_config(2, 20, manytor_b200.UR5_ARM, 0, action_low=-90, action_high=91)
    a = _host_action_sample(ur5, 2)
    assert len(a) == 2 and len(a[0]) == 6 and all(-90 <= v <= 90 for row in a for v in row)
