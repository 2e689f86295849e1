#!/bin/bash
# dev tool (run under gpurun, one GPU): the evidence files of a round.
#   bench line, ncu launch list of the same command, `ncu --set full` of the dominant kernels,
#   the secondary configs, `ncu --set full` of the symmetry-function kernels.
# usage: tools/profile_round.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_f64.json 2> gpurun_out/${tag}_bench.err || exit 1
cat gpurun_out/${tag}_bench_f64.json
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --check-atoms 8 > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --check-atoms 8 \
    > gpurun_out/${tag}_ncu_list.log 2>&1
python tools/agg_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_by_kernel.txt 2>&1
head -20 gpurun_out/${tag}_launches_by_kernel.txt
for prec in high medium; do
python tools/eamz_sweep.py --child --precision $prec --steps 3 > gpurun_out/${tag}_plain_$prec.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_eamz' -s 6 -c 2 -f \
    -o gpurun_out/${tag}_eamz_$prec python tools/eamz_sweep.py --child --precision $prec --steps 3 \
    > gpurun_out/${tag}_ncu_$prec.log 2>&1
done
python tools/build_breakdown.py 0.3 > gpurun_out/${tag}_build.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/${tag}_build_launches.csv python tools/build_breakdown.py 0.3 > gpurun_out/${tag}_ncu_build_list.log 2>&1
python tools/agg_launches.py gpurun_out/${tag}_build_launches.csv > gpurun_out/${tag}_build_launches_by_kernel.txt 2>&1
python tools/build_breakdown.py 0.0 >> gpurun_out/${tag}_build.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/${tag}_build_exact_launches.csv python tools/build_breakdown.py 0.0 > gpurun_out/${tag}_ncu_build_exact_list.log 2>&1
python tools/agg_launches.py gpurun_out/${tag}_build_exact_launches.csv > gpurun_out/${tag}_build_exact_launches_by_kernel.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_nbr_tile' -s 4 -c 1 -f \
    -o gpurun_out/${tag}_nbr python tools/build_breakdown.py 0.3 > gpurun_out/${tag}_ncu_nbr.log 2>&1
python tools/bench_configs.py > gpurun_out/${tag}_other_configs.jsonl 2> gpurun_out/${tag}_other_configs.err
cat gpurun_out/${tag}_other_configs.jsonl
for prec in high medium; do
python tools/sf_profile_run.py $prec > gpurun_out/${tag}_sf_plain_$prec.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_sf_backward|k_sf_forward' -s 8 -c 2 -f \
    -o gpurun_out/${tag}_sf_$prec python tools/sf_profile_run.py $prec > gpurun_out/${tag}_ncu_sf_$prec.log 2>&1
done
ls -la gpurun_out/${tag}_*
