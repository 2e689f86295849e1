"""A/B of the lane-split EAM kernels on the 1 M-atom headline (dev tool, run under gpurun):
per-kernel CUDA-event times for lanes per atom x virial form x precision, one process per
setting.  usage: python tools/eamz_sweep.py [--cells 63] [--libs libA.so,libB.so]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(args):
    import torch
    import bench
    from tensoralloy_b200 import _lib
    from tensoralloy_b200.nn.eam.potentials import get_potential
    pot = get_potential('zjw04')
    model = _lib.EamModel(_lib.EAM_ALLOY, 1, [pot.rho('Ni')], [pot.phi('NiNi')],
                          [pot.embed('Ni')])
    prec = _lib.PRECISION_HIGH if args.precision == 'high' else _lib.PRECISION_MEDIUM
    pos, cell = bench.make_lattice(args.cells)
    d_pos = torch.from_numpy(pos).cuda()
    nbr = _lib.NeighborList()
    if args.skin > 0:
        nbr.set_skin(args.skin)
    nbr.build(d_pos, None, cell, [1, 1, 1], bench.RC)
    n = len(pos)
    e = torch.zeros(1, dtype=torch.float64, device='cuda')
    f = torch.zeros((n, 3), dtype=torch.float64, device='cuda')
    v = torch.zeros(9, dtype=torch.float64, device='cuda')
    for _ in range(3):
        nbr.update(d_pos)
        model.eval(nbr, prec, energy=e, forces=f, virial=v)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        nbr.update(d_pos)
        model.eval(nbr, prec, energy=e, forces=f, virial=v)
    ev1.record()
    torch.cuda.synchronize()
    ms, calls = _lib.profile_read()
    out = {"L": os.environ.get('TAB_EAMZ_L'), "vir": os.environ.get('TAB_EAMZ_VIR'),
           "lib": os.path.basename(os.environ.get('TAB200_LIB', 'libtab200.so')),
           "precision": args.precision, "skin": args.skin,
           "step_ms": ev0.elapsed_time(ev1) / args.steps,
           "rho_ms": ms[0], "spread_ms": ms[1], "force_ms": ms[2], "reduce_ms": ms[3],
           "nij": nbr.sizes()[0], "energy": e.item(), "f_l2": float(torch.linalg.norm(f)),
           "virial_trace": float(v[0] + v[4] + v[8])}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cells', type=int, default=63)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--precision', default='high')
    ap.add_argument('--skin', type=float, default=0.0)
    ap.add_argument('--child', action='store_true')
    ap.add_argument('--libs', default='')
    ap.add_argument('--lanes', default='0,1,2,4,8')
    ap.add_argument('--virs', default='0')
    ap.add_argument('--precisions', default='high,medium')
    args = ap.parse_args()
    if args.child:
        return child(args)
    libs = [x for x in args.libs.split(',') if x] or ['libtab200.so']
    for lib in libs:
        for prec in args.precisions.split(','):
            for L in args.lanes.split(','):
                for vir in args.virs.split(','):
                    if L == '0' and vir != args.virs.split(',')[0]:
                        continue
                    env = dict(os.environ, TAB_EAMZ_L=L, TAB_EAMZ_VIR=vir,
                               TAB200_LIB=os.path.join(ROOT, 'tensoralloy_b200', 'csrc', lib))
                    r = subprocess.run([sys.executable, __file__, '--child', '--cells',
                                        str(args.cells), '--steps', str(args.steps),
                                        '--precision', prec, '--skin', str(args.skin)],
                                       env=env, capture_output=True, text=True, timeout=600)
                    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else \
                        json.dumps({"error": r.stderr[-400:], "L": L, "vir": vir, "lib": lib})
                    print(line, flush=True)


if __name__ == '__main__':
    main()
