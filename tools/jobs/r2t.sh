python -m pytest tests/test_gpu_round2.py tests/test_gpu_dropin.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2t_bench.json")); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["dram_frac"]); print(d["e2e"]); print(d["config1_single_env"])
PY
