#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dd_check.py 12 > gpurun_out/r02zy_dd_check.log 2>&1
echo "dd_check rc=$?"; grep "md-cycle\|FAIL" gpurun_out/r02zy_dd_check.log | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > gpurun_out/r02zy_bench_n2.json 2> gpurun_out/r02zy_bench_n2.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02zy_bench_n2.json'))
print('n2 value %.4g ms %.4f resident %.4f e2e %.3f rebuilds %d'%(d['value'],d['ms_per_step'],d['resident']['ms_per_step'],d['e2e']['ms_per_step'],d['config']['rebuilds_in_timed_steps']), d['check']['ok'], d['check']['energy'], d['check']['f_l2'], d['check']['max_disp_since_build'])
PY
grep -v "^W\|^\[W\|OMP\|^\*\|^$\|NCCL" gpurun_out/r02zy_bench_n2.err | tail -5
