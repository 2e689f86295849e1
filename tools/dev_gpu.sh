set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; tail -3 gpurun_out/bench_r1b.err; cat gpurun_out/bench_r1b.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision medium > gpurun_out/bench_r1b_f32.json 2>> gpurun_out/bench_r1b.err; cat gpurun_out/bench_r1b_f32.json
