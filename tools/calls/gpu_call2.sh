#!/bin/bash
# round 2, GPU call 2: tests, prefetch-depth sweep, ncu of the lane-split kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_eam_fast_gpu.py -x -q > gpurun_out/r02b_fast_tests.log 2>&1
echo "fast tests rc=$?" >> gpurun_out/r02b_fast_tests.log
tail -15 gpurun_out/r02b_fast_tests.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02b_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r02b_gpu_suite.log
tail -8 gpurun_out/r02b_gpu_suite.log
timeout 600 python tools/eamz_sweep.py --lanes 0,1,2,4 > gpurun_out/r02b_sweep.jsonl 2> gpurun_out/r02b_sweep.err
timeout 600 python tools/eamz_sweep.py --libs libtab200_pf1.so,libtab200_pf3.so,libtab200_mb4.so,libtab200_ldplain.so,libtab200_t256.so --lanes 1,2 >> gpurun_out/r02b_sweep.jsonl 2>> gpurun_out/r02b_sweep.err
timeout 200 python tools/eamz_sweep.py --skin 0.3 --lanes 1 >> gpurun_out/r02b_sweep.jsonl 2>> gpurun_out/r02b_sweep.err
cat gpurun_out/r02b_sweep.jsonl
export TAB_EAMZ_L=1
for prec in high medium; do
python tools/eamz_sweep.py --child --precision $prec --steps 3 > gpurun_out/r02b_plain_$prec.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_eamz' -s 6 -c 4 -f \
    -o gpurun_out/r02b_eamz_$prec python tools/eamz_sweep.py --child --precision $prec --steps 3 \
    > gpurun_out/r02b_ncu_$prec.log 2>&1
tail -3 gpurun_out/r02b_ncu_$prec.log
done
ls -la gpurun_out/r02b_*
