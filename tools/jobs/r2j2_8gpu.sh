TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2j2_bench_n8.json 2> gpurun_out/r2j2_bench_n8.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2j2_bench_n8.json")); print(d["n_gpus"], d["value"], d["ms_per_step"], d["repeats"]["best_ms_per_step"], d["roofline"]["kernel_ms_per_launch"], json.dumps(d["e2e"])[:500])
for k,v in d["modes"].items(): print("   ", k, round(v["us_per_step_median"],2), f'{v["env_steps_per_s"]:.3e}')
PY
