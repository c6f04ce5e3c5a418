python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
MT_HOST_ZEROCOPY=1 python bench.py --steps 20 --warmup 5 --no-cpu --repeats 3 > gpurun_out/r2k_bench_zc.json 2> gpurun_out/r2k_bench_zc.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err
python - <<'PY'
import json
for f in ("gpurun_out/r2k_bench.json","gpurun_out/r2k_bench_zc.json"):
    d=json.load(open(f)); print(f, d["value"], d["ms_per_step"], d["repeats"]["best_ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["e2e"]["value"], d["e2e"]["frac_of_d2h_ceiling"])
    for k,v in d["modes"].items(): print("   ", k, round(v["us_per_step_median"],2), round(v["us_per_step_best"],2))
d=json.load(open("gpurun_out/r2k_bench.json")); print(d["cpu_baseline"]); print(d["config1_single_env"])
print(open("gpurun_out/r2k_bench_ref.json").read()[:400])
PY
