#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_final_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r02_final_gpu_suite.log
grep -E "passed|failed|^FAILED|^ERROR|suite rc" gpurun_out/r02_final_gpu_suite.log | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
