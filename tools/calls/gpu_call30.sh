#!/bin/bash
mkdir -p gpurun_out
for skin in 0.0 0.3; do python tools/build_breakdown.py $skin 2>&1 | tail -1; done | tee gpurun_out/r02y_build.log
timeout 900 python -m pytest tests/test_nbr_gpu.py tests/test_eam_fast_gpu.py tests/test_eam_gpu.py tests/test_domain_gpu.py tests/test_finite_temperature_gpu.py -m gpu -q -x 2>&1 | tail -3
for n in 1024 8192 65536; do timeout 300 python tools/td_heads_bench.py $n 2>&1 | tail -2 | cut -c1-200; done | tee gpurun_out/r02y_td_heads_sizes.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02y_build_launches.csv python tools/build_breakdown.py 0.3 > gpurun_out/r02y_ncu.log 2>&1
python tools/agg_launches.py gpurun_out/r02y_build_launches.csv 2>&1 | head -6
