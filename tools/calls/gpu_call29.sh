#!/bin/bash
# 8 GPUs: bench N = 8 with and without the CUDA-graph step, N = 4 and N = 2 after the pinned-buffer fix
mkdir -p gpurun_out
run() { # N tag extra...
  N=$1; tag=$2; shift; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 20 --warmup 5 --rebuild-profile "$@" > gpurun_out/r02x_bench_$tag.json 2> gpurun_out/r02x_bench_$tag.err
  echo "rc=$? $tag"; python - <<PY
import json
d=json.load(open('gpurun_out/r02x_bench_$tag.json'))
print('$tag value %.4g ms %.4f resident %.4f e2e %.3f rebuilds %d'%(d['value'],d['ms_per_step'],d['resident']['ms_per_step'],d['e2e']['ms_per_step'],d['config']['rebuilds_in_timed_steps']), {k:round(v,2) for k,v in d['config']['rebuild_profile_ms'].items()}, d['check']['ok'], d['check']['energy'], d['check']['f_l2'])
PY
}
run 8 n8
run 8 n8_nograph --no-graph
run 4 n4
run 2 n2
