#!/bin/bash
# dev tool (run under `gpurun --gpus 8`): the 8- and 4-GPU bench lines of the headline config,
# strong and weak scaling, with and without the CUDA-graph step -> gpurun_out/scale_<tag>.json
mkdir -p gpurun_out
run() { # N tag extra...
  N=$1; tag=$2; shift; shift
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$N bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 5 "$@" > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  echo "rc=$? $tag: $(wc -l < gpurun_out/scale_$tag.json) stdout lines"; python -c "
import json
for l in open('gpurun_out/scale_$tag.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', 'ms %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], 'e2e %.3f'%d['e2e']['ms_per_step'], d['roofline']['kernel_ms'])
"
}
run 8 n8
run 4 n4
run 8 n8w --scaling weak
run 8 n8_nograph --no-graph
