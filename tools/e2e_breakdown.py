"""Wall-time breakdown of the e2e EAM call (dev tool): H2D, list build, eval, D2H."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tensoralloy_b200 import _lib
from tensoralloy_b200.nn.eam.potentials import get_potential

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 63
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
pot = get_potential('zjw04')
model = _lib.EamModel(_lib.EAM_ALLOY, 1, [pot.rho('Ni')], [pot.phi('NiNi')], [pot.embed('Ni')])
pos, cell = bench.make_lattice(cells)
n = len(pos)
h_pos = torch.from_numpy(pos).pin_memory()
d_pos = torch.empty((n, 3), dtype=torch.float64, device='cuda')
d_e = torch.zeros(1, dtype=torch.float64, device='cuda')
d_f = torch.zeros((n, 3), dtype=torch.float64, device='cuda')
d_v = torch.zeros(9, dtype=torch.float64, device='cuda')
h_f = torch.zeros((n, 3), dtype=torch.float64).pin_memory()
nbr = _lib.NeighborList()
sync = torch.cuda.synchronize
for rep in range(reps):
    sync(); t0 = time.perf_counter()
    d_pos.copy_(h_pos, non_blocking=True); sync(); t1 = time.perf_counter()
    nbr.build(d_pos, None, cell, [1, 1, 1], bench.RC); sync(); t2 = time.perf_counter()
    model.eval(nbr, _lib.PRECISION_HIGH, energy=d_e, forces=d_f, virial=d_v); sync(); t3 = time.perf_counter()
    h_f.copy_(d_f, non_blocking=True); sync(); t4 = time.perf_counter()
    print(f"mode={os.environ.get('TAB_NBR_MODE','default')} n={n} h2d {1e3*(t1-t0):.3f} build {1e3*(t2-t1):.3f} "
          f"eval {1e3*(t3-t2):.3f} d2h {1e3*(t4-t3):.3f} ms", flush=True)
