#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/td_heads_bench.py 65536 2>&1 | tail -3 | tee gpurun_out/r02o_td_heads.jsonl
timeout 900 python -m pytest tests/test_finite_temperature_gpu.py tests/test_batch_gpu.py tests/test_training_gpu.py -m gpu -q -x > gpurun_out/r02o_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r02o_tests.log
