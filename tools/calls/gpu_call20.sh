#!/bin/bash
mkdir -p gpurun_out
for A in 16 8; do
  echo "tile $A"; TAB_TD_TILE=$A timeout 300 python tools/td_heads_bench.py 65536 2>&1 | tail -2 | cut -c1-200
done | tee gpurun_out/r02q_td_heads_ring.log
timeout 600 python -m pytest tests/test_finite_temperature_gpu.py -m gpu -q -x 2>&1 | tail -3
