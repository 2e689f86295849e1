#!/bin/bash
# 8 GPUs, final multi-GPU evidence: multi-GPU tests, md-cycle checks at 4 / 8 ranks, bench N = 8, 4, 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_domain_multigpu.py -m gpu -q -x > gpurun_out/r02zb_multigpu_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02zb_multigpu_tests.log
for N in 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N tools/dd_check.py > gpurun_out/r02zb_dd_check_n$N.log 2>&1
echo "dd_check N=$N rc=$?"; grep "md-cycle" gpurun_out/r02zb_dd_check_n$N.log | head -2
done
run() { # N tag extra...
  N=$1; tag=$2; shift; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 20 --warmup 5 --rebuild-profile "$@" > gpurun_out/r02zb_bench_$tag.json 2> gpurun_out/r02zb_bench_$tag.err
  echo "rc=$? $tag"; python - <<PY
import json
d=json.load(open('gpurun_out/r02zb_bench_$tag.json'))
print('$tag value %.4g ms %.4f resident %.4f e2e %.3f rebuilds %d'%(d['value'],d['ms_per_step'],d['resident']['ms_per_step'],d['e2e']['ms_per_step'],d['config']['rebuilds_in_timed_steps']), {k:round(v,2) for k,v in d['config']['rebuild_profile_ms'].items()}, d['check']['ok'], d['check']['energy'], d['check']['f_l2'])
PY
}
run 8 n8
run 4 n4
run 2 n2
