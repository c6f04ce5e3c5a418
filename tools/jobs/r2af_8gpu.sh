TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2af_bench_n8.json 2> gpurun_out/r2af_bench_n8.err
MT_STATS_PEERS=0 $TR --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 5 --repeats 6 > gpurun_out/r2af_bench_n8_nccl.json 2>> gpurun_out/r2af_bench_n8.err
python -m pytest tests/test_gpu_round2.py -m gpu -q -k "peer_memory" 2>&1 | tail -2
python - <<'PY'
import json
for f in ("gpurun_out/r2af_bench_n8.json","gpurun_out/r2af_bench_n8_nccl.json"):
    d=json.load(open(f)); print(f, d["value"], d["ms_per_step"], d["repeats"]["best_ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["e2e"]["value"], d["e2e"]["host_step_mode"], d["config"]["parallelism"][-60:])
PY
