mkdir -p gpurun_out
python -m pytest tests/test_eam_gpu.py tests/test_host_logic.py -m gpu -x -q 2>&1 | tail -3
for z in 1 0; do TAB_HOST_ZEROCOPY=$z python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 10 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('zerocopy=$z step %.4f'%d['ms_per_step'], 'e2e %.3f'%d['e2e']['ms_per_step'])
"; done
