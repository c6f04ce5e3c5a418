# final verification of round 2 on the committed tree (one GPU)
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2bh_tests.log; tail -3 gpurun_out/r2bh_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2bh_smoke.log 2>&1; tail -2 gpurun_out/r2bh_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2bh_bench.json 2> gpurun_out/r2bh_bench.err; cut -c1-300 gpurun_out/r2bh_bench.json
B="python bench.py --steps 20 --warmup 5 --no-cpu --repeats 2 --burn-in 100 --clock-warm-s 0 --clock-probe-s 0 --no-graph"
MT_HOST_ZEROCOPY=0 $B > gpurun_out/r2bh_bench_plain.json 2> gpurun_out/r2bh_bench_plain.err && MT_HOST_ZEROCOPY=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2bh_launches.csv $B > gpurun_out/r2bh_bench_under_ncu.json 2> gpurun_out/r2bh_ncu_launch.err
wc -l gpurun_out/r2bh_launches.csv
