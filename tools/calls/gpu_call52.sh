#!/bin/bash
mkdir -p gpurun_out
TAB_DD_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > gpurun_out/r02zy3_bench_n2.json 2> gpurun_out/r02zy3_bench_n2.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02zy3_bench_n2.json'))
print('n2 ms %.4f resident %.4f'%(d['ms_per_step'],d['resident']['ms_per_step']))
PY
grep "md_step" gpurun_out/r02zy3_bench_n2.err | awk '{ if ($5+0 > 1.5) print }' | head -20
grep -c "md_step" gpurun_out/r02zy3_bench_n2.err
