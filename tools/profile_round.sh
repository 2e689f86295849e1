#!/bin/bash
# dev tool (run under gpurun): bench line, ncu launch list of the same command and one
# `ncu --set full` capture of the dominant kernels.  usage: tools/profile_round.sh <tag>
tag=${1:-r01x}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_f64.json 2> gpurun_out/${tag}_bench.err || exit 1
cat gpurun_out/${tag}_bench_f64.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_eam_force|k_eam_rho|k_nbr_tile' \
    -s 6 -c 3 -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_*
