#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_size_gpu.py tests/test_domain_multigpu.py -m gpu -q -x > gpurun_out/r02z_fullsize_multigpu_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r02z_fullsize_multigpu_tests.log
