#!/bin/bash
# last check of the final code on one GPU: full suite, smoke, bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02zz_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r02zz_gpu_suite.log
grep -E "passed|failed|^FAILED|^ERROR|suite rc" gpurun_out/r02zz_gpu_suite.log | tail -10
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r02zz_bench_f64.json 2> gpurun_out/r02zz_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02zz_bench_f64.json'))
print('value %.4g ms %.4f resident %.4f e2e %.3f medium %.4g'%(d['value'],d['ms_per_step'],d['resident']['ms_per_step'],d['e2e']['ms_per_step'],d['extra']['medium']['value']), d['clocks'], d['check']['ok'], d['roofline']['traffic'])
PY
