#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02d_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r02d_gpu_suite.log
tail -12 gpurun_out/r02d_gpu_suite.log
timeout 400 python tools/eamz_sweep.py --lanes 1 > gpurun_out/r02d_sweep.jsonl 2> gpurun_out/r02d_sweep.err
timeout 400 python tools/eamz_sweep.py --libs libtab200_swapring.so,libtab200_rhonotab.so --lanes 1 >> gpurun_out/r02d_sweep.jsonl 2>> gpurun_out/r02d_sweep.err
timeout 300 python tools/eamz_sweep.py --skin 0.3 --lanes 1 >> gpurun_out/r02d_sweep.jsonl 2>> gpurun_out/r02d_sweep.err
TAB_NBR_SORT_ROWS=0 timeout 300 python tools/eamz_sweep.py --skin 0.3 --lanes 1 >> gpurun_out/r02d_sweep.jsonl 2>> gpurun_out/r02d_sweep.err
cat gpurun_out/r02d_sweep.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
echo "bench rc=$?"
cat gpurun_out/r02d_bench.json; tail -5 gpurun_out/r02d_bench.err
