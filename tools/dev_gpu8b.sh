mkdir -p gpurun_out
run() { # N tag extra...
  N=$1; tag=$2; shift; shift
  timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$N bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 5 "$@" > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  echo "rc=$? $tag"; python -c "
import json
for l in open('gpurun_out/scale_$tag.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', 'ms %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], 'e2e %.3f'%d['e2e']['ms_per_step'], d['roofline']['kernel_ms'])
"
}
run 8 r01k_n8
run 8 r01k_n8w --scaling weak
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 tools/train_check.py 128 2>&1 | grep -i "world\|error" | tail -2
