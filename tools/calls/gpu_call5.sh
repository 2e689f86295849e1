#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eam_fast_gpu.py tests/test_domain_gpu.py tests/test_eam_gpu.py -q -x > gpurun_out/r02e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02e_tests.log
tail -5 gpurun_out/r02e_tests.log
timeout 400 python tools/eamz_sweep.py --lanes 1 > gpurun_out/r02e_sweep.jsonl 2> gpurun_out/r02e_sweep.err
timeout 300 python tools/eamz_sweep.py --skin 0.3 --lanes 1 >> gpurun_out/r02e_sweep.jsonl 2>> gpurun_out/r02e_sweep.err
TAB_NBR_SORT_ROWS=0 timeout 300 python tools/eamz_sweep.py --skin 0.3 --lanes 1 >> gpurun_out/r02e_sweep.jsonl 2>> gpurun_out/r02e_sweep.err
cat gpurun_out/r02e_sweep.jsonl
python tools/build_breakdown.py 0.3 > gpurun_out/r02e_build.log 2>&1 && python tools/build_breakdown.py 0.0 >> gpurun_out/r02e_build.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02e_build_launches.csv python tools/build_breakdown.py 0.3 > gpurun_out/r02e_build_ncu.log 2>&1
cat gpurun_out/r02e_build.log
python tools/agg_launches.py gpurun_out/r02e_build_launches.csv 2>/dev/null | head -30
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
echo "bench rc=$?"
cat gpurun_out/r02e_bench.json; tail -5 gpurun_out/r02e_bench.err
