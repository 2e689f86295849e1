#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02c_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r02c_gpu_suite.log
tail -8 gpurun_out/r02c_gpu_suite.log
timeout 600 python tools/eamz_sweep.py --lanes 0,1,2 > gpurun_out/r02c_sweep.jsonl 2> gpurun_out/r02c_sweep.err
timeout 600 python tools/eamz_sweep.py --libs libtab200_rhotab.so,libtab200_pf2.so --lanes 1 >> gpurun_out/r02c_sweep.jsonl 2>> gpurun_out/r02c_sweep.err
timeout 200 python tools/eamz_sweep.py --skin 0.3 --lanes 1 >> gpurun_out/r02c_sweep.jsonl 2>> gpurun_out/r02c_sweep.err
cat gpurun_out/r02c_sweep.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
echo "bench rc=$?"
cat gpurun_out/r02c_bench.json; tail -20 gpurun_out/r02c_bench.err
