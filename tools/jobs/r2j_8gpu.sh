# 8-GPU job: D2H ceiling, bench at N=8 (both host-step variants), C-ABI collective tests
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2j_topo.txt 2>&1
$TR --master-port 29511 tools/microbench/d2h_ceiling.py > gpurun_out/r2j_ceiling_plain.json 2> gpurun_out/r2j_ceiling.err
$TR --master-port 29512 tools/microbench/d2h_ceiling.py --bind > gpurun_out/r2j_ceiling_bind.json 2>> gpurun_out/r2j_ceiling.err
$TR --master-port 29513 tools/microbench/d2h_ceiling.py --bind --h2d > gpurun_out/r2j_ceiling_bind_h2d.json 2>> gpurun_out/r2j_ceiling.err
python tools/microbench/d2h_ceiling.py > gpurun_out/r2j_ceiling_1gpu.json 2>> gpurun_out/r2j_ceiling.err
$TR --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err
MT_HOST_ZEROCOPY=1 $TR --master-port 29515 bench.py --gpus 8 --steps 20 --warmup 5 --repeats 2 > gpurun_out/r2j_bench_n8_zerocopy.json 2> gpurun_out/r2j_bench_n8_zc.err
python -m pytest tests/test_gpu_round2.py -m gpu -q -k "allreduce or c_host or multi_step or zero_copy" 2>&1 | tail -6 > gpurun_out/r2j_tests.log
cat gpurun_out/r2j_ceiling_*.json; tail -4 gpurun_out/r2j_tests.log
python - <<'PY'
import json
for f in ("gpurun_out/r2j_bench_n8.json","gpurun_out/r2j_bench_n8_zerocopy.json"):
    try:
        d=json.load(open(f)); print(f, d["value"], d["ms_per_step"], json.dumps(d["e2e"])[:600])
    except Exception as ex: print(f, "unreadable", ex)
PY
