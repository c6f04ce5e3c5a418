"""Multi-GPU plumbing: environments shard contiguously across ranks, one process
per GPU, NO per-step collective (envs never interact, manytor.py:115-122); the
only exchange is one all-reduce (NCCL over NVLink on GPUs, gloo in CPU tests) of
the MT_STATS_WORDS episode-statistics vector at the end of a rollout.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from ._lib import STATS_FIELDS


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of [0, n_total): returns (first global env id, count) of `rank`.
    Global ids key the RNG, so results do not depend on the shard count."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(int(n_total), world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Join the process group torchrun described (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*);
    returns (rank, world, local_rank).  No-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def bind_to_gpu_numa(local_rank: int) -> Optional[list]:
    """Pin this process to the CPU cores next to its GPU (NVML's ideal affinity) so that pinned host
    buffers and the H2D/D2H copies of the host-buffer step stay on the GPU's NUMA node.  Returns the
    cores, or None when NVML / affinity control is unavailable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cores = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the per-shard statistics vector over all ranks (in place) -- the single
    collective of a multi-GPU rollout."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


class StatsReducer:
    """The one collective of a multi-GPU rollout as ONE small kernel of the library over NVLink peer memory
    (`mt_stats_allreduce_peers`): every rank's 64-byte statistics go straight into every peer's buffer, a flag per
    peer says when.  torch's symmetric memory only maps the buffers (plumbing).  Where that is unavailable -- a
    single process, a CPU group, an older torch -- `reduce` falls back to `all_reduce` (NCCL / gloo), and says so
    in `self.path`."""

    def __init__(self, device=None, group=None, use_peers: bool = True):
        self.group = group
        self.path = "local"
        self._peers = self._buf = self._hdl = None
        if not (dist.is_initialized() and dist.get_world_size(group) > 1):
            return
        self.path = "all_reduce"
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if not use_peers or device is None or torch.device(device).type != "cuda" or os.environ.get("MT_STATS_PEERS", "1") == "0":
            return
        try:
            import ctypes as C
            import torch.distributed._symmetric_memory as symm
            from . import _lib
            words = int(_lib.load().mt_stats_peer_buffer_bytes(self.world)) // 8
            self._buf = symm.empty(words, dtype=torch.int64, device=device)
            self._buf.zero_()
            self._hdl = symm.rendezvous(self._buf, group if group is not None else dist.group.WORLD)
            ptrs = [int(p) for p in self._hdl.buffer_ptrs]
            if len(ptrs) != self.world or any(p == 0 for p in ptrs):
                raise RuntimeError("symmetric memory returned no peer pointers")
            self._peers = torch.tensor(ptrs, dtype=torch.int64, device=device)
            torch.cuda.synchronize(device)
            self.path = "peer memory kernel (mt_stats_allreduce_peers over NVLink, buffers mapped by torch symmetric memory)"
        except Exception as ex:                     # noqa: BLE001 -- any failure here only selects the fallback
            self._peers = None
            self.path = f"all_reduce (peer-memory path unavailable: {type(ex).__name__}: {str(ex)[:120]})"
        # the choice must be the same on every rank (a rank on the fallback would leave the others waiting for its
        # flag): one MIN over a success bit; this collective is also the barrier after which every buffer is zeroed
        ok = torch.tensor([1 if self._peers is not None else 0], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0 and self._peers is not None:
            self._peers = None
            self.path = "all_reduce (peer-memory path unavailable on another rank)"

    def reduce(self, env) -> torch.Tensor:
        """Global sum of `env`'s statistics (MT_STATS_WORDS int64 on its device), asynchronous on the current stream."""
        if self._peers is None:
            return allreduce_stats(env.stats_tensor(), self.group)
        import ctypes as C
        from . import _lib
        out = torch.empty((len(STATS_FIELDS),), dtype=torch.int64, device=env.device)
        stream = C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)
        _lib.check(env._lib.mt_stats_allreduce_peers(env._h, C.c_void_p(self._peers.data_ptr()), self.rank, self.world,
                                                      C.c_void_p(out.data_ptr()), stream))
        return out


def stats_dict(stats: torch.Tensor) -> dict:
    v = stats.detach().cpu().tolist()
    return {k: int(v[i]) for i, k in enumerate(STATS_FIELDS)}


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (device timings are reported as the slowest rank)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
