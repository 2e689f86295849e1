mkdir -p gpurun_out
python -m pytest tests/test_nbr_gpu.py tests/test_domain_gpu.py tests/test_eam_gpu.py tests/test_atomic_gpu.py -m gpu -x -q 2>&1 | tail -8
for mode in tile; do TAB_NBR_MODE=$mode python tools/e2e_breakdown.py 63 3 | tail -1; done
TAB_NBR_MODE=tile ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_build_tile.csv python tools/e2e_breakdown.py 63 2 > /dev/null 2>&1
