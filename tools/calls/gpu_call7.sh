#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra --rebuild-profile > gpurun_out/r02g_bench_n2_prof.json 2> gpurun_out/r02g_bench_n2_prof.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench_n2_prof.json'))
print(d['ms_per_step'], d['resident'], d['config']['rebuilds_in_timed_steps'], d['config']['rebuild_profile_ms'], d['check']['ok'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29554 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench_n2.json'))
print(d['ms_per_step'], d['resident'], d['config']['rebuilds_in_timed_steps'], d['check']['ok'])
PY
grep -v "^W\|^\[W" gpurun_out/r02g_bench_n2.err | tail -5
