mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --e2e-steps 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo rc=$?
wc -c gpurun_out/bench_n2.json gpurun_out/bench_n2.err
tail -30 gpurun_out/bench_n2.err
cut -c1-600 gpurun_out/bench_n2.json
