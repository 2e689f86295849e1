#!/bin/bash
for w in 2 8; do python tools/dd_build_breakdown.py $w 0.3 2>&1 | tail -1; done
for skin in 0.0 0.3; do python tools/build_breakdown.py $skin 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_nbr_gpu.py tests/test_domain_gpu.py tests/test_eam_fast_gpu.py -m gpu -q -x 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g ms %.4f e2e %.3f'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step']), d['check']['ok'])"
