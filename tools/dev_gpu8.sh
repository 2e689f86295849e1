mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m pytest tests/test_domain_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for N in 2 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 --scaling weak > gpurun_out/scale_n${N}w.json 2> gpurun_out/scale_n${N}w.err
done
for f in gpurun_out/scale_n*.json; do python - $f <<'PY'
import sys,json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms']; print(sys.argv[1], d['n_gpus'], d['scaling'], 'value %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], {a:round(b,3) for a,b in k.items()}, 'e2e ms %.3f'%d['e2e']['ms_per_step'])
PY
done
tail -3 gpurun_out/scale_n8.err
