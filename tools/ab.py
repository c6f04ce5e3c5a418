#!/usr/bin/env python
"""A/B timing of several builds of libmanytor_b200.so in ONE process on ONE GPU, interleaved
round by round so that clocks, temperature and the box are the same for all of them.
  python tools/ab.py [--lg 20] [--x 10] [--arm ref|ur5] [--rounds 6] [--steps 400] A.so B.so[@keep_mb] [C.so ...]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from manytor_b200 import BatchedEnvs, REFERENCE_ARM, UR5_ARM

ap = argparse.ArgumentParser()
ap.add_argument("libs", nargs="+")
ap.add_argument("--lg", type=int, default=20)
ap.add_argument("--x", type=int, default=10)
ap.add_argument("--arm", default="ref")
ap.add_argument("--rounds", type=int, default=6)
ap.add_argument("--steps", type=int, default=400)
ap.add_argument("--modes", default="step,rand,noobs")
ap.add_argument("--fk-mode", type=int, default=0)
ap.add_argument("--isolate", type=int, default=0, metavar="REPS",
                help="one PROCESS per library, REPS passes over the list (A B C A B C ...): handles that share a "
                     "process also share the L2, and the evict_last state lines of the idle ones stay resident")
args = ap.parse_args()

if args.isolate:
    import re, subprocess
    base = [sys.executable, os.path.abspath(__file__), "--lg", str(args.lg), "--x", str(args.x), "--arm", args.arm,
            "--rounds", str(args.rounds), "--steps", str(args.steps), "--modes", args.modes, "--fk-mode", str(args.fk_mode)]
    acc = {}
    for rep in range(args.isolate):
        for lib in args.libs:
            out = subprocess.run(base + [lib], capture_output=True, text=True).stdout
            for line in out.splitlines():
                m = re.match(r"(.{32}) (\S+)\s+median\s+([0-9.]+)", line)
                if m:
                    acc.setdefault((m.group(1), lib), []).append(float(m.group(3)))
    print(f"N = 2^{args.lg} envs, x = {args.x}, arm = {args.arm}, {args.steps} steps x {args.rounds} rounds, "
          f"{args.isolate} isolated processes per library (medians of each)")
    for (mode, lib), v in acc.items():
        print(f"{mode} {os.path.basename(lib):20s} " + "  ".join(f"{t:7.2f}" for t in v) + f"   best {min(v):7.2f} us/step")
    sys.exit(0)

# "lib.so@48" = create that handle with MT_L2_KEEP_MB=48; "lib.so@MT_WARPS_PER_BLOCK=14" sets any variable
specs = [(p.split("@") + [None])[:2] for p in args.libs]
paths = [os.path.abspath(p) for p, _ in specs]
names = [os.path.basename(p) + (("@" + k) if k else "") for p, k in specs]
n, K = 1 << args.lg, args.steps
arm = UR5_ARM if args.arm == "ur5" else REFERENCE_ARM
J = arm.n_joints
envs = []
for p, (_, keep) in zip(paths, specs):
    for k in ("MT_L2_KEEP_MB", "MT_WARPS_PER_BLOCK", "MT_TAIL_RANKS", "MT_TAIL_RATE"):
        os.environ.pop(k, None)
    for kv in (keep.split(",") if keep else []):     # "48" = MT_L2_KEEP_MB=48; "K=V" sets any variable
        k, _, v = kv.rpartition("=")
        os.environ[k or "MT_L2_KEEP_MB"] = v
    e = BatchedEnvs(n, args.x, arm=arm, device=0, auto_reset=True, horizon=1000, seed=1, lib_path=p, fk_mode=args.fk_mode)
    e.reset()
    e.rollout_random(1000, write_obs=False)
    envs.append(e)
acts = [torch.randint(-180, 180, (n, J), device="cuda").float() for _ in range(8)]
torch.cuda.synchronize()
t0 = time.perf_counter()
while time.perf_counter() - t0 < 1.5:
    for e in envs:
        e.rollout_random(100)
    torch.cuda.synchronize()

def run(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / K

graphs = {}
def step_graph(e):          # K steps as one CUDA graph: Python cannot issue a launch every ~45 us
    if id(e) not in graphs:
        e.step(acts[0]); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(K):
                e.step(acts[i & 7])
        g.replay(); torch.cuda.synchronize()
        graphs[id(e)] = g
    graphs[id(e)].replay()

all_modes = {
    "step": ("step(actions from HBM), obs", step_graph),
    "rand": ("rollout_random, obs", lambda e: e.rollout_random(K)),
    "noobs": ("rollout_random, no obs", lambda e: e.rollout_random(K, write_obs=False)),
}
modes = {all_modes[m][0]: all_modes[m][1] for m in args.modes.split(",")}
res = {m: [[] for _ in envs] for m in modes}
for r in range(args.rounds):
    for m, fn in modes.items():
        for i, e in enumerate(envs):
            res[m][i].append(run(lambda: fn(e)))
print(f"N = 2^{args.lg} envs, x = {args.x}, arm = {args.arm}, {K} steps x {args.rounds} rounds")
for m in modes:
    for i, p in enumerate(paths):
        v = sorted(res[m][i])
        print(f"{m:32s} {names[i]:20s} median {v[len(v)//2]:8.2f} us/step  min {v[0]:8.2f}  max {v[-1]:8.2f}")
