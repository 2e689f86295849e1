#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02j_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r02j_gpu_suite.log
grep -E "float32|passed|failed|^FAILED" gpurun_out/r02j_gpu_suite.log | tail -40
