python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2bi_bench_n8.json 2> gpurun_out/r2bi_bench_n8.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2bi_bench_n8.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("host_step_mode"), d["config"]["parallelism"][-110:])
PY
tail -2 gpurun_out/r2bi_bench_n8.err
