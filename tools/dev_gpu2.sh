set -x
mkdir -p gpurun_out
nvidia-smi -L
for N in 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r1c_n$N.json 2> gpurun_out/bench_r1c_n$N.err; tail -5 gpurun_out/bench_r1c_n$N.err; cat gpurun_out/bench_r1c_n$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --scaling weak > gpurun_out/bench_r1c_n${N}w.json 2> gpurun_out/bench_r1c_n${N}w.err; tail -5 gpurun_out/bench_r1c_n${N}w.err; cat gpurun_out/bench_r1c_n${N}w.json
done
