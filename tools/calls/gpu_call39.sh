#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02zc_bench_f64.json 2> gpurun_out/r02zc_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02zc_bench_f64.json'))
print('value %.4g ms %.4f resident %.4f e2e %.3f'%(d['value'],d['ms_per_step'],d['resident']['ms_per_step'],d['e2e']['ms_per_step']), d['clocks'], d['check']['ok'])
PY
tail -3 gpurun_out/r02zc_bench.err
