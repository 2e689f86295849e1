#!/bin/bash
# ncu --set full of the EAM pair kernels on the lists the bench uses (0.3 A skin); build A/B for the skin flag
mkdir -p gpurun_out
for prec in high medium; do
python tools/eamz_sweep.py --child --precision $prec --skin 0.3 --steps 3 > gpurun_out/r02u_plain_$prec.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_eamz' -s 6 -c 2 -f \
    -o gpurun_out/r02u_eamz_skin_$prec python tools/eamz_sweep.py --child --precision $prec --skin 0.3 --steps 3 \
    > gpurun_out/r02u_ncu_$prec.log 2>&1
done
for skin in 0.0 0.001 0.05 0.1 0.2 0.3; do python tools/build_breakdown.py $skin 2>&1 | tail -1; done | tee gpurun_out/r02u_build_vs_skin.log
ls -la gpurun_out/r02u*
