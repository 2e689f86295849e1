#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dd_check.py 12 > gpurun_out/r02i_dd_check.log 2>&1
echo "dd_check rc=$?" >> gpurun_out/r02i_dd_check.log
grep -v "^W\|^\[W\|OMP\|^\*\|^$" gpurun_out/r02i_dd_check.log | tail -12
bash tools/gpu_call7.sh
