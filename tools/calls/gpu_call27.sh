#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02v_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r02v_gpu_suite.log
grep -E "passed|failed|^FAILED|^ERROR|suite rc" gpurun_out/r02v_gpu_suite.log | tail -30
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
