#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,clocks.max.sm,clocks.max.mem,power.draw,temperature.gpu,ecc.mode.current --format=csv
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02zz2_bench_f64.json 2> gpurun_out/r02zz2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02zz2_bench_f64.json')); m=d['extra']['medium']
print('value %.4g ms %.4f resident %.4f e2e %.3f'%(d['value'],d['ms_per_step'],d['resident']['ms_per_step'],d['e2e']['ms_per_step']), d['roofline']['kernel_ms'])
print('medium', m['ms_per_step'], m['resident_ms_per_step'], m['roofline']['kernel_ms'])
PY
