#!/bin/bash
# 2 GPUs: decomposition checks (peer + NCCL paths, MD cycle with migration) and bench --gpus 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_domain_multigpu.py -q -x > gpurun_out/r02f_multigpu_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02f_multigpu_tests.log
tail -30 gpurun_out/r02f_multigpu_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dd_check.py 12 > gpurun_out/r02f_dd_check.log 2>&1
echo "dd_check rc=$?" >> gpurun_out/r02f_dd_check.log
grep -v "^W\|^\[W" gpurun_out/r02f_dd_check.log | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err
echo "bench rc=$?"
cat gpurun_out/r02f_bench_n2.json; grep -v "^W\|^\[W" gpurun_out/r02f_bench_n2.err | tail -15
