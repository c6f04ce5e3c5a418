python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 20 --warmup 5 --repeats 4 > gpurun_out/r2bf_bench_n4.json 2> gpurun_out/r2bf_bench_n4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2bf_bench_n4.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["parallelism"][-110:])
PY
tail -3 gpurun_out/r2bf_bench_n4.err
