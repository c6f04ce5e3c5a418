python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; tail -2 gpurun_out/r2q_smoke.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_bench_ref_n2.json 2> gpurun_out/r2q_bench_ref_n2.err
python -m pytest tests/test_gpu_round2.py -m gpu -q -k "allreduce or c_host or policy_loop" 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2q_bench_n2.json")); print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["frac_of_d2h_ceiling"], d["gpu_launches"])
print(open("gpurun_out/r2q_bench_ref_n2.json").read()[:300])
PY
