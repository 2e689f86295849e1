#!/bin/bash
# round 2, GPU call 1: new tests, full suite, lane-split sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_eam_fast_gpu.py -x -q > gpurun_out/r02a_fast_tests.log 2>&1
echo "fast tests rc=$?" >> gpurun_out/r02a_fast_tests.log
tail -15 gpurun_out/r02a_fast_tests.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02a_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r02a_gpu_suite.log
tail -8 gpurun_out/r02a_gpu_suite.log
timeout 900 python tools/eamz_sweep.py > gpurun_out/r02a_sweep.jsonl 2> gpurun_out/r02a_sweep.err
timeout 300 python tools/eamz_sweep.py --libs libtab200_mb4.so,libtab200_mb6.so --lanes 4,8 --virs 1 --precisions high >> gpurun_out/r02a_sweep.jsonl 2>> gpurun_out/r02a_sweep.err
timeout 300 python tools/eamz_sweep.py --skin 0.3 --lanes 0,4 --virs 1 >> gpurun_out/r02a_sweep.jsonl 2>> gpurun_out/r02a_sweep.err
cat gpurun_out/r02a_sweep.jsonl
