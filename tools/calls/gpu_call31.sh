#!/bin/bash
mkdir -p gpurun_out
for skin in 0.0 0.3; do python tools/build_breakdown.py $skin 2>&1 | tail -1; done | tee gpurun_out/r02y2_build.log
timeout 900 python -m pytest tests/test_nbr_gpu.py tests/test_eam_fast_gpu.py tests/test_eam_gpu.py tests/test_domain_gpu.py tests/test_neighbor_sizes_gpu.py tests/test_sqlite.py -m gpu -q -x 2>&1 | tail -3
