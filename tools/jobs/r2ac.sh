python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err
python tools/run_config.py rollout 20 > gpurun_out/r2ac_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/r2ac_prof_rollout python tools/run_config.py rollout 20 > gpurun_out/r2ac_ncu2.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ac_bench.json")); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["dram_frac"], d["e2e"]["value"], d["config1_single_env"]["us_per_step"])
for k,v in d["modes"].items(): print("   ", k, round(v["us_per_step_median"],2), f'{v["env_steps_per_s"]:.3e}')
PY
