#!/usr/bin/env python
"""Benchmark of the ManyTor step loop on B200 (contract: see the task README / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input: ONE
fused-kernel launch advancing every env of the shard by one env-step, actions
read from HBM, observations/reward/done written to HBM, on-device auto-reset with
objective refresh (BASELINE.json configs[2]: 4-joint arm, 2^20 envs per GPU, x=10,
1000-step horizon).  N>1 runs under torchrun, one rank per GPU, envs sharded
(weak scaling, 2^20 per GPU, configs[3]); the only collective is one all-reduce of
the episode statistics at the end of the timed region.

`--impl reference` times the CPU restatement of the reference loop (oracle/, a
vectorised numpy port pinned to the reference's golden traces; the reference
itself is pure Python and cannot travel to the GPU box) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 1 << 20
OBJ = 10
HORIZON = 1000
SEED = 20201
METRIC = "env-steps/s"
WORKLOAD = "4-joint reference arm, 2^20 envs per GPU, x=10 objectives, 1000-step horizon, on-device auto-reset + objective refresh (BASELINE configs[2]/[3])"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------
# clocks during the timed region
# ----------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_ms: int = 200):
        self.gpu, self.proc, self.lines, self.period_ms = gpu_index, None, [], int(period_ms)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", str(self.period_ms)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)
        return {"sm_mhz": sm_sorted[len(sm_sorted) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------
# CPU legs (the ONLY place bench.py touches oracle/)
# ----------------------------------------------------------------------------
def _oracle_worker_init(n_envs, x, seed):
    global _W
    import numpy as np
    from oracle import OracleEnvs
    rng = np.random.RandomState(seed)
    env = OracleEnvs(n_envs, x)
    v = rng.normal(size=(n_envs, x, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    v[..., 2] = np.abs(v[..., 2])
    env.reset(points=v * (51.3 * rng.uniform(size=(n_envs, x, 1)) ** (1 / 3)))
    _W = (env, rng)


def _oracle_worker_step(_):
    env, rng = _W
    import numpy as np
    r = env.step(rng.randint(-180, 180, size=(env.n, 4)))
    if r.done.any():                                   # refresh like test_single.py:20-21,32
        v = rng.normal(size=(env.n, env.x, 3))
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        v[..., 2] = np.abs(v[..., 2])
        env.reset(mask=r.done, points=v * (51.3 * rng.uniform(size=(env.n, env.x, 1)) ** (1 / 3)))
    return int(r.reward.sum())


def cpu_baseline_single_core(budget_s: float = 12.0) -> dict:
    """Vectorised fp64 numpy oracle on ONE core, N=4096, x=10 ("host numpy" of the metric string),
    plus the scalar per-env loop port (test_multi.py shape) as a second figure."""
    import numpy as np
    n = 4096
    _oracle_worker_init(n, OBJ, 1)
    _oracle_worker_step(0)
    t0, steps = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s:
        _oracle_worker_step(0)
        steps += 1
    dt = time.perf_counter() - t0
    from oracle.scalar_port import multienv_loop
    t1 = time.perf_counter()
    es, _ = multienv_loop(64, OBJ, 60, seed=0)       # Multienv((8,8), 10)-sized, ~4 s
    dts = time.perf_counter() - t1
    return {"value": n * steps / dt, "unit": METRIC, "cores": 1, "kind": "port",
            "sample": f"oracle/manytor_oracle.py (vectorised fp64 numpy restatement of manytor.py:175-260), "
                      f"{n} envs x {steps} steps, x={OBJ}, random integer actions, {dt:.1f}s on 1 core",
            "scalar_loop_value": es / dts,
            "scalar_loop_sample": f"oracle/scalar_port.py (per-env Python loop shaped like test_multi.py / "
                                  f"manytor.py:115-122), 64 envs x 60 steps in {dts:.1f}s on 1 core"}


def run_reference_arm(args) -> None:
    """`--impl reference`: the CPU port of the reference loop on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, min(cores, 64))
    n = 4096
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    ctx = mp.get_context("fork")
    pools = [ctx.Pool(1, initializer=_oracle_worker_init, initargs=(n, OBJ, 100 + i)) for i in range(procs)]

    def step_all():
        rs = [p.apply_async(_oracle_worker_step, (0,)) for p in pools]
        return [r.get() for r in rs]

    for _ in range(max(args.warmup, 1)):
        step_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_all()
    dt = time.perf_counter() - t0
    for p in pools:
        p.terminate()
    value = procs * n * args.steps / dt
    sample = (f"oracle/manytor_oracle.py (vectorised fp64 numpy port of the reference step, pinned to the "
              f"reference's golden traces), {procs} processes x {n} envs x {args.steps} steps, x={OBJ}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU arm: each step = one env-step of a bounded sample "
                       f"({procs} x {n} envs) of the workload"},
            "cpu_baseline": {"value": value, "unit": METRIC, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _RESULT_LINE.append(json.dumps(line))


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def run_gpu(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist
    import manytor_b200
    from manytor_b200 import BatchedEnvs, distributed as mtd

    rank, world, local = mtd.init_from_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    manytor_b200.load_library()
    cores = mtd.bind_to_gpu_numa(local) if world > 1 else None      # keep pinned buffers on the GPU's NUMA node

    n = ENVS_PER_GPU
    base, _ = rank * n, n
    env = BatchedEnvs(n, OBJ, device=local, env_id_base=base, horizon=HORIZON, auto_reset=True, seed=SEED)
    env.reset()
    # burn-in to the steady state of the episode process (resets every ~550 steps per env)
    env.rollout_random(args.burn_in, write_obs=False)
    K, W = args.steps, max(args.warmup, 3)
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED + rank)
    n_act = min(K + W, 64)
    actions = [torch.randint(-180, 180, (n, 4), device=dev, generator=gen).float() for _ in range(n_act)]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    stream = torch.cuda.current_stream(dev)

    def timed(fn, k, w):
        for i in range(w):
            fn(i)
        torch.cuda.synchronize(); barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = env.launch_count
        e0.record(stream)
        for i in range(k):
            fn(w + i)
        e1.record(stream)
        torch.cuda.synchronize(); barrier(); torch.cuda.synchronize()
        return e0.elapsed_time(e1), env.launch_count - l0

    # ---- headline: K x mt_step, actions from HBM, obs written (309 B/env-step) ----------
    sampler = ClockSampler(local, args.clock_period_ms) if rank == 0 and args.clock_period_ms > 0 else None
    stats_holder = {}

    def step_fn(i):
        env.step(actions[i % n_act])

    # bring clocks / power state to steady load before timing (untimed, same kernel)
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < args.clock_warm_s:
        for i in range(200):
            step_fn(i)
        torch.cuda.synchronize()
    # The K timed steps are K launches of the step kernel either way; by default they are submitted
    # as ONE CUDA graph (captured untimed, each node its own action buffer), which is how an RL loop
    # that graphs policy + env drives it and removes the ~3 us per-launch gap of stream launches.
    graph = None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(stream)
            with torch.cuda.stream(side):
                for i in range(3):
                    step_fn(i)
            stream.wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for i in range(K):
                    step_fn(W + i)
        except Exception as ex:
            print(f"bench: CUDA graph capture failed ({ex}); timing stream launches", file=sys.stderr)
            graph = None
    if sampler:
        sampler.start()
    for i in range(W):
        step_fn(i)
    if graph is not None:
        graph.replay()                            # untimed: first replay uploads the graph
    torch.cuda.synchronize(); barrier(); torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    l0 = env.launch_count
    e0.record(stream)
    if graph is not None:
        graph.replay()                            # exactly K steps
    else:
        for i in range(K):
            step_fn(W + i)
    e1.record(stream)
    st = env.stats_tensor()                       # end-of-rollout statistics ...
    mtd.allreduce_stats(st)                       # ... one NCCL all-reduce (config 4)
    e2.record(stream)
    torch.cuda.synchronize(); barrier(); torch.cuda.synchronize()
    # my kernels inside the timed region: K step kernels (graph nodes or stream launches) + 1 stats kernel
    launches_timed = (K + 1) if graph is not None else env.launch_count - l0
    timed_ms_local = e0.elapsed_time(e2)
    # K steps last only a few ms, far below nvidia-smi's sampling period: keep the same step
    # loop running (untimed) so the clock sampler sees the load the timed region ran under
    t_cont = time.perf_counter()
    while sampler and time.perf_counter() - t_cont < args.clock_probe_s:
        for i in range(200):
            step_fn(i)
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["note"] = (f"sampled every {args.clock_period_ms} ms from just before the warm-up steps, through the {timed_ms_local:.1f} ms "
                          f"timed region, and over a {args.clock_probe_s:.1f} s untimed continuation of the same step loop")
    launches = launches_timed
    ms_total = mtd.max_over_ranks(e0.elapsed_time(e2), dev)
    ms_kernels = mtd.max_over_ranks(e0.elapsed_time(e1), dev)
    stats = mtd.stats_dict(st)
    value = world * n * K / (ms_total * 1e-3)
    B = env.bytes_per_env_step(True, True)
    kernel_ms = ms_kernels / K
    peak, peak_src = peaks()
    achieved = B * n / (kernel_ms * 1e-3) / 1e9

    # ---- other modes (short), for the roofline discussion -------------------------------
    modes = {}
    KM = min(K, 500)
    ms, _ = timed(lambda i: env.rollout_random(1, write_obs=True), KM, 3)
    modes["rollout_random_in_kernel_actions"] = {"env_steps_per_s": n * KM / (ms * 1e-3), "bytes_per_env_step": env.bytes_per_env_step(False, True)}
    ms, _ = timed(lambda i: env.rollout_random(1, write_obs=False), KM, 3)
    modes["rollout_random_no_obs_write"] = {"env_steps_per_s": n * KM / (ms * 1e-3), "bytes_per_env_step": env.bytes_per_env_step(False, False)}
    # the same steps as individual stream launches from Python (no graph)
    ms, _ = timed(step_fn, KM, 3)
    modes["step_hbm_actions_stream_launches"] = {"env_steps_per_s": n * KM / (ms * 1e-3), "bytes_per_env_step": B}
    for m in modes.values():
        if "error" in m:
            continue
        m["hbm_gbs"] = m["env_steps_per_s"] * m["bytes_per_env_step"] / 1e9
        m["frac_of_peak"] = m["hbm_gbs"] / peak

    # ---- e2e: host buffers through the public API (H2D actions, D2H obs/reward/done) ----
    act_pinned = env.pinned("actions", (n, 4), np.float32)
    act_pinned[:] = np.random.RandomState(rank).randint(-180, 180, size=(n, 4)).astype(np.float32)
    e2e_steps = max(3, min(K, 10))
    for _ in range(2):
        env.step_host(act_pinned)
    torch.cuda.synchronize(); barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        obs_h, rew_h, done_h = env.step_host(act_pinned)      # synchronous: results are in host memory
    torch.cuda.synchronize(); barrier()
    e2e_s = mtd.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_value = world * n * e2e_steps / e2e_s
    h2d = n * 4 * 4
    d2h = n * (3 * OBJ * 4 + 4 + 1)

    if rank == 0:
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu = cpu_baseline_single_core() if world == 1 and not args.no_cpu else None
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n, "n_joints": 4, "n_obj": OBJ, "horizon": HORIZON,
                       "terminate_on_ground": False, "auto_reset": True, "actions": "uniform integer degrees in [-180,180), read from HBM [N][4] fp32",
                       "submission": "one K-node CUDA graph" if graph is not None else "K stream launches",
                       "burn_in_steps": args.burn_in,
                       "l2": f"working set {(n * (B + 8 * 4)) / 2**20:.0f} MiB per step > 126 MiB L2 (inputs larger than L2, no flush needed)",
                       "parallelism": f"env-sharded x{world}, no per-step collective, 1 stats all-reduce"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "mt::step_kernel<0,10,false,true>",
                         "algorithmic_bytes_per_env_step": B, "kernel_ms_per_launch": kernel_ms,
                         "envs_per_launch": n,
                         # the algorithmic count includes the 48 B/env-step of per-env state that the kernel keeps in
                         # L2 from launch to launch, so `frac` can pass 1; this is the DRAM side on its own
                         "dram_achieved": (traffic / kernel_ms / 1e6) if traffic else None,
                         "dram_frac": (traffic / kernel_ms / 1e6 / peak) if traffic else None},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "BatchedEnvs.step_host -> mt_step_host (pinned host buffers, 16 chunks over 4 streams)",
                    "cpu_binding": f"{len(cores)} cores next to the GPU (NVML affinity)" if cores else "none"},
            "gpu_launches": launches,
            "clocks": clocks,
            "modes": modes,
            "episode_stats": stats,
        }
        _RESULT_LINE.append(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_RESULT_LINE = []      # the one JSON line, printed by main() after stdout has been restored


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--clock-warm-s", type=float, default=1.0, help="untimed seconds of the step loop before warm-up")
    ap.add_argument("--clock-probe-s", type=float, default=1.0, help="untimed continuation for the clock sampler")
    ap.add_argument("--clock-period-ms", type=int, default=200, help="nvidia-smi sampling period (0 = no sampler)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="time K stream launches instead of one K-node CUDA graph")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--burn-in", type=int, default=HORIZON,
                    help="untimed random-action steps before warm-up so episodes reach their steady state")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: anything a library writes to file descriptor 1 on the
    # way (NCCL prints its version there when the box sets NCCL_DEBUG) is sent to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_gpu(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if _RESULT_LINE:
        print(_RESULT_LINE[-1], flush=True)


if __name__ == "__main__":
    main()
