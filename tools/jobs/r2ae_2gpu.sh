python -m pytest tests/test_gpu_round2.py -m gpu -q -k "peer_memory or allreduce or c_host" 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 4 > gpurun_out/r2ae_bench_n2.json 2> gpurun_out/r2ae_bench_n2.err
MT_STATS_PEERS=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 4 > gpurun_out/r2ae_bench_n2_nccl.json 2>> gpurun_out/r2ae_bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2ae_bench_n2.json","gpurun_out/r2ae_bench_n2_nccl.json"):
    d=json.load(open(f)); print(f, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["config"]["parallelism"][-140:])
PY
tail -5 gpurun_out/r2ae_bench_n2.err
