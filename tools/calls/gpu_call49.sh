#!/bin/bash
timeout 600 python -m pytest tests/test_eam_fast_gpu.py -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02zy_bench_f64.json 2>/dev/null; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02zy_bench_f64.json')); m=d['extra']['medium']
print('value %.4g ms %.4f resident %.4f e2e %.3f rebuilds %d'%(d['value'],d['ms_per_step'],d['resident']['ms_per_step'],d['e2e']['ms_per_step'],d['config']['rebuilds_in_timed_steps']), d['check']['ok'], d['check']['max_dF'], d['check']['reused_vs_fresh_lists'])
print('medium', m['value'], m['ms_per_step'], m['rebuilds_in_timed_steps'])
PY
