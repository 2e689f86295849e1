#!/bin/bash
# final 1-GPU evidence: suite, smoke, bench line, launch list, ncu --set full of k_nbr_tile, TD heads sizes
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02z_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r02z_gpu_suite.log
grep -E "passed|failed|^FAILED|^ERROR|suite rc" gpurun_out/r02z_gpu_suite.log | tail -10
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r02z_bench_f64.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --check-atoms 8 > gpurun_out/r02z_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --check-atoms 8 > gpurun_out/r02z_ncu_list.log 2>&1
python tools/agg_launches.py gpurun_out/r02z_launches.csv > gpurun_out/r02z_launches_by_kernel.txt 2>&1; head -8 gpurun_out/r02z_launches_by_kernel.txt
python tools/build_breakdown.py 0.3 > gpurun_out/r02z_build.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_nbr_tile' -s 4 -c 1 -f -o gpurun_out/r02z_nbr python tools/build_breakdown.py 0.3 > gpurun_out/r02z_ncu_nbr.log 2>&1
python tools/build_breakdown.py 0.0 >> gpurun_out/r02z_build.log 2>&1; cat gpurun_out/r02z_build.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_reference_arm.json 2> gpurun_out/r02z_reference_arm.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02z_reference_arm.json
