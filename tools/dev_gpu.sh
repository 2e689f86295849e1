mkdir -p gpurun_out
for v in base estrin estrin_t64 estrin_t256 unroll2 estrin_u2 minb5 estrin_minb5; do
TAB200_LIB=$PWD/build/variants/lib_$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms']; print('$v', 'step %.4f'%d['ms_per_step'], 'rho %.4f force %.4f'%(k['rho_pass'],k['force_pass']))
"; done
