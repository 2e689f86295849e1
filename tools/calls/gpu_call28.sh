#!/bin/bash
mkdir -p gpurun_out
for G in 4 2 1; do
  echo "split $G"; TAB_TD_SPLIT=$G timeout 300 python tools/td_heads_bench.py 65536 2>&1 | tail -2 | cut -c1-200
done | tee gpurun_out/r02w_td_heads_smem_ops.log
timeout 600 python -m pytest tests/test_finite_temperature_gpu.py tests/test_batch_gpu.py -m gpu -q -x 2>&1 | tail -3
