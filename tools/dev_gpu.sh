mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_eam_|k_count|k_fill" -s 4 -c 4 -o gpurun_out/prof_r1d python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
