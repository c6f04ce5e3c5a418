python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_shape.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2f_tests.log
tools/microbench/fma_mix_bin > gpurun_out/r2f_fma_mix.txt 2>&1
python tools/ab.py --isolate 2 --rounds 5 --steps 400 build/variants/r2opt1.so manytor_b200/lib/libmanytor_b200.so > gpurun_out/r2f_ab.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -6 gpurun_out/r2f_tests.log; cat gpurun_out/r2f_fma_mix.txt gpurun_out/r2f_ab.txt; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print(d['value'], d['ms_per_step'], d['repeats']['best_ms_per_step'])
for k,v in d['modes'].items(): print(k, v['us_per_step_median'], v['us_per_step_best'], v['env_steps_per_s'])
PY
