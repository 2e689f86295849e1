#!/bin/bash
# N = 2: multi-GPU tests, md-cycle check (migration + rebuilds against the oracle), bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_domain_multigpu.py -m gpu -q -x > gpurun_out/r02l_multigpu_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02l_multigpu_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dd_check.py 12 > gpurun_out/r02l_dd_check.log 2>&1
echo "dd_check rc=$?"
grep -v "^W\|^\[W\|OMP\|^\*\|^$" gpurun_out/r02l_dd_check.log | tail -8
sed -i 's/r02k_bench/r02l_bench/g' tools/gpu_call13.sh
bash tools/gpu_call13.sh 2
