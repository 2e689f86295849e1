#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 --rebuild-profile > gpurun_out/r02k_bench_n$N.json 2> gpurun_out/r02k_bench_n$N.err
echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02k_bench_n$N.json'))
print('value',d['value'],'ms',d['ms_per_step'],'resident',d['resident']['ms_per_step'],'rebuilds',d['config']['rebuilds_in_timed_steps'],d['config']['rebuild_profile_ms'])
print('check',d['check'])
print('e2e',d['e2e']['ms_per_step'],'medium',d['extra']['medium']['ms_per_step'],d['extra']['medium']['resident_ms_per_step'])
print(d['config']['parallelism'])
PY
grep -v "^W\|^\[W\|OMP\|^\*\|^$\|NCCL" gpurun_out/r02k_bench_n$N.err | tail -8
