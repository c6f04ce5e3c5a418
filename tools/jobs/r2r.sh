python -m pytest tests/test_gpu_round2.py -m gpu -q -k "multi_step" 2>&1 | tail -4
python tools/big_shard_check.py 2>&1 | tail -3
python tools/bench_obj_counts.py 2>&1 | tail -6
