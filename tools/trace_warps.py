#!/usr/bin/env python
"""Where does a step launch lose time at its ends?  Needs a -DMT_TRACE build of the library:
  nvcc <build.py flags> -DMT_TRACE -o tools/ab/trace.so manytor_b200/csrc/mt_api.cu
  python tools/trace_warps.py tools/ab/trace.so [lg_envs [raw.npy]]
Prints, for the last of a few hundred back-to-back steps: how blocks were placed on SMs, the spread of
warp start and finish times, tiles per warp, and the share of warp-time between first start and last
finish in which a warp slot was already empty."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from manytor_b200 import BatchedEnvs

path = os.path.abspath(sys.argv[1])
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = 1 << lg
env = BatchedEnvs(n, 10, device=0, auto_reset=True, horizon=1000, seed=1, lib_path=path)
env.reset()
env.rollout_random(300, write_obs=False)
acts = [torch.randint(-180, 180, (n, 4), device="cuda").float() for _ in range(8)]
for i in range(200):
    env.step(acts[i & 7])
torch.cuda.synchronize()
lib = C.CDLL(path)
buf = np.zeros((8192, 4), dtype=np.uint64)
assert lib.mt_debug_trace(buf.ctypes.data_as(C.c_void_p)) == 0
tr = buf[buf[:, 1] > 0].astype(np.int64)
if len(sys.argv) > 3:                       # raw rows (start ns, end ns, SM, tiles), warp w = block * warps_per_block + warp
    np.save(sys.argv[3], tr)
t0, t1, sm, tiles = tr[:, 0], tr[:, 1], tr[:, 2], tr[:, 3]
base = t0.min()
span = t1.max() - base
print(f"{len(tr)} warps, {len(np.unique(sm))} SMs, launch span {span/1e3:.2f} us (first warp start -> last warp end)")
per_sm = np.bincount(sm)
print(f"warps per SM: min {per_sm[per_sm>0].min()} max {per_sm.max()}")
q = lambda v: " ".join(f"{np.percentile(v, p)/1e3:6.2f}" for p in (0, 10, 50, 90, 100))
print(f"warp start after launch start (us), p0/10/50/90/100: {q(t0 - base)}")
print(f"warp end   after launch start (us), p0/10/50/90/100: {q(t1 - base)}")
print(f"tiles per warp: " + ", ".join(f"{k}: {int(v)}" for k, v in enumerate(np.bincount(tiles)) if v))
for k in np.unique(tiles):
    m = tiles == k
    print(f"  warps with {k} tiles: busy time p10/50/90 {np.percentile((t1-t0)[m],10)/1e3:.2f} {np.percentile((t1-t0)[m],50)/1e3:.2f} "
          f"{np.percentile((t1-t0)[m],90)/1e3:.2f} us")
print(f"warp-slot residency over the span: {np.sum(t1 - t0) / (len(tr) * span) * 100:.1f}%  "
      f"(idle before start {np.sum(t0 - base)/(len(tr)*span)*100:.1f}%, after end {np.sum(t1.max() - t1)/(len(tr)*span)*100:.1f}%)")
# per-SM: last finish minus first finish
spread = [t1[sm == s].max() - t1[sm == s].min() for s in np.unique(sm)]
print(f"per-SM spread between first and last warp finish: median {np.median(spread)/1e3:.2f} us, max {np.max(spread)/1e3:.2f} us")
ends = np.array([t1[sm == s].max() - base for s in np.unique(sm)])
print(f"per-SM finish time: min {ends.min()/1e3:.2f} median {np.median(ends)/1e3:.2f} max {ends.max()/1e3:.2f} us")
# does the SM that starts late finish late?  (a launch's blocks become resident as the previous launch's blocks exit)
sms = np.unique(sm)
starts = np.array([t0[sm == s].min() - base for s in sms])
print(f"per-SM first warp start: p50 {np.median(starts)/1e3:.2f} p90 {np.percentile(starts,90)/1e3:.2f} max {starts.max()/1e3:.2f} us; "
      f"correlation(start, finish) over SMs = {np.corrcoef(starts, ends)[0,1]:.2f}")
order = np.argsort(-ends)[:8]
print("latest SMs (sm: start -> finish us, tiles done): " +
      ", ".join(f"{int(sms[i])}: {starts[i]/1e3:.2f} -> {ends[i]/1e3:.2f}, {int(tiles[sm == sms[i]].sum())}" for i in order))
per_sm_tiles = np.array([tiles[sm == s].sum() for s in sms])
print(f"tiles per SM: min {per_sm_tiles.min()} median {int(np.median(per_sm_tiles))} max {per_sm_tiles.max()}; "
      f"busy span (finish - start) per SM: p10 {np.percentile(ends-starts,10)/1e3:.2f} p50 {np.median(ends-starts)/1e3:.2f} "
      f"p90 {np.percentile(ends-starts,90)/1e3:.2f} max {(ends-starts).max()/1e3:.2f} us")
