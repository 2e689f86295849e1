#!/bin/bash
mkdir -p gpurun_out
for w in 2 4 8; do python tools/dd_build_breakdown.py $w 0.3 2>&1 | tail -1; done | tee gpurun_out/r02z_dd_build.log
for w in 8; do for sd in 1 2; do echo "subdiv $sd: $(TAB_NBR_SUBDIV=$sd python tools/dd_build_breakdown.py $w 0.3 2>&1 | tail -1)"; done; done | tee -a gpurun_out/r02z_dd_build.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_dd_build_launches.csv python tools/dd_build_breakdown.py 8 0.3 > gpurun_out/r02z_dd_ncu.log 2>&1
python tools/agg_launches.py gpurun_out/r02z_dd_build_launches.csv 2>&1 | head -24
