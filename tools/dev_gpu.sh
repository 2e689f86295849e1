set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -3 gpurun_out/bench_r1c.err; cat gpurun_out/bench_r1c.json
