mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/time_breakdown.py 2>&1 | tail -2
python tools/bench_configs.py 2>&1 | tee gpurun_out/configs_r1.jsonl | cut -c1-330
