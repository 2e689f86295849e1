#!/bin/bash
mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus $N --steps 20 --warmup 5 --rebuild-profile > gpurun_out/r02zd_bench_n8.json 2> gpurun_out/r02zd_bench_n8.err
echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02zd_bench_n8.json'))
print('n8 value %.4g ms %.4f resident %.4f e2e %.3f rebuilds %d'%(d['value'],d['ms_per_step'],d['resident']['ms_per_step'],d['e2e']['ms_per_step'],d['config']['rebuilds_in_timed_steps']), {k:round(v,2) for k,v in d['config']['rebuild_profile_ms'].items()}, d['check']['ok'], d['check']['energy'], d['clocks'])
PY
