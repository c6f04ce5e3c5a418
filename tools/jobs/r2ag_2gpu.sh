python -m pytest tests/test_gpu_round2.py -m gpu -q -k "peer_memory or allreduce" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 3 > gpurun_out/r2ag_bench_n2.json 2> gpurun_out/r2ag_bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ag_bench_n2.json")); print(d["value"], d["ms_per_step"], d["episode_stats"]["env_steps"], d["config"]["parallelism"][-90:])
PY
