#!/bin/bash
# 8 GPUs: md-cycle check with migration at 4 and 8 ranks (peer-memory migration), bench lines N = 8, 4
mkdir -p gpurun_out
for N in 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N tools/dd_check.py > gpurun_out/r02t_dd_check_n$N.log 2>&1
echo "dd_check N=$N rc=$?"
grep -v "^W\|^\[W\|OMP\|^\*\|^$\|NCCL" gpurun_out/r02t_dd_check_n$N.log | grep "md-cycle\|FAIL\|Error\|error" | head -10
done
sed -i 's/r02[a-z]_bench/r02t_bench/g' tools/gpu_call13.sh
bash tools/gpu_call13.sh 8
bash tools/gpu_call13.sh 4
