#!/bin/bash
# 1 GPU: half-cell neighbour build A/B (TAB_NBR_SUBDIV), neighbour / EAM tests, new data-format tests
mkdir -p gpurun_out
for sd in 1 2; do
  for skin in 0.0 0.3; do
    echo "subdiv $sd: $(TAB_NBR_SUBDIV=$sd timeout 300 python tools/build_breakdown.py $skin 2>&1 | tail -1)"
  done
  for prec in high medium; do
    TAB_NBR_SUBDIV=$sd timeout 300 python tools/eamz_sweep.py --child --precision $prec --skin 0.3 --steps 10 2>&1 | tail -1 | cut -c1-330
  done
done | tee gpurun_out/r02m_subdiv_ab.log
timeout 1200 python -m pytest tests/test_nbr_gpu.py tests/test_eam_fast_gpu.py tests/test_eam_gpu.py tests/test_domain_gpu.py tests/test_wire_format_gpu.py tests/test_sqlite.py tests/test_neighbor_sizes_gpu.py -m gpu -q -x > gpurun_out/r02m_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02m_tests.log
