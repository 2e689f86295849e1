#!/bin/bash
mkdir -p gpurun_out
run() { # N tag
  N=$1; tag=$2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N bench.py --gpus $N --steps 20 --warmup 5 --rebuild-profile > gpurun_out/r02zz_bench_$tag.json 2> gpurun_out/r02zz_bench_$tag.err
  echo "rc=$? $tag"; python - <<PY
import json
d=json.load(open('gpurun_out/r02zz_bench_$tag.json'))
print('$tag value %.4g ms %.4f resident %.4f e2e %.3f'%(d['value'],d['ms_per_step'],d['resident']['ms_per_step'],d['e2e']['ms_per_step']), {k:round(v,2) for k,v in d['config']['rebuild_profile_ms'].items()}, d['check']['ok'], d['check']['energy'])
PY
}
run 4 n4
run 2 n2
