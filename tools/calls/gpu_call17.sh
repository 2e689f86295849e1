#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_finite_temperature_gpu.py tests/test_batch_gpu.py tests/test_nbr_gpu.py tests/test_eam_fast_gpu.py tests/test_domain_gpu.py tests/test_training_gpu.py -m gpu -q -x > gpurun_out/r02n_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r02n_tests.log
timeout 300 python tools/td_heads_bench.py 65536 2>&1 | tail -2 | tee gpurun_out/r02n_td_heads.jsonl
for skin in 0.0 0.3; do python tools/build_breakdown.py $skin 2>&1 | tail -1; done
TAB_NBR_SUBDIV=2 timeout 300 python tools/eamz_sweep.py --child --precision high --skin 0.3 --steps 10 2>&1 | tail -1 | cut -c1-330
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02n_build_launches.csv python tools/build_breakdown.py 0.3 > gpurun_out/r02n_ncu.log 2>&1
python tools/agg_launches.py gpurun_out/r02n_build_launches.csv 2>&1 | head -30
