#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02h_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r02h_gpu_suite.log
tail -12 gpurun_out/r02h_gpu_suite.log
timeout 400 python tools/eamz_sweep.py --lanes 1 > gpurun_out/r02h_sweep.jsonl 2> gpurun_out/r02h_sweep.err
timeout 400 python tools/eamz_sweep.py --libs libtab200_mb4.so,libtab200_mb6.so,libtab200_simple64.so --lanes 1 --precisions high >> gpurun_out/r02h_sweep.jsonl 2>> gpurun_out/r02h_sweep.err
timeout 300 python tools/eamz_sweep.py --skin 0.3 --lanes 1 >> gpurun_out/r02h_sweep.jsonl 2>> gpurun_out/r02h_sweep.err
cat gpurun_out/r02h_sweep.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
echo "bench rc=$?"
cat gpurun_out/r02h_bench.json; tail -5 gpurun_out/r02h_bench.err
