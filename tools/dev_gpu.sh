mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 5 --precision medium 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('f32 step %.4f'%d['ms_per_step'], 'e2e %.3f'%d['e2e']['ms_per_step'], d['roofline']['kernel_ms'])
"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; tail -2 gpurun_out/bench_r1h.err; cut -c1-400 gpurun_out/bench_r1h.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_r1h.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_eam_force|k_eam_rho|k_nbr_tile" -c 3 -o gpurun_out/prof_r1h -f python tools/e2e_breakdown.py 63 1 > gpurun_out/ncu_r1h_full.log 2>&1
tail -2 gpurun_out/ncu_r1h_full.log
