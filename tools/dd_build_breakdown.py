"""List build of ONE rank of a `world`-way slab decomposition of the 1 M-atom headline, on one GPU
(dev tool; run it under `ncu --metrics gpu__time_duration.sum` for the launch list):
python tools/dd_build_breakdown.py [world] [skin]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np     # noqa: E402
import torch           # noqa: E402
import bench           # noqa: E402
from tensoralloy_b200 import _lib                                   # noqa: E402
from tensoralloy_b200.domain import SlabLayout, SlabRank            # noqa: E402
from tensoralloy_b200.nn.eam.potentials import get_potential        # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
skin = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
pos, cell = bench.make_lattice(63)
lx, ly, lz = cell[0, 0], cell[1, 1], cell[2, 2]
pos[:, 0] = np.mod(pos[:, 0], lx)
pot = get_potential('zjw04')
model = _lib.EamModel(_lib.EAM_ALLOY, 1, [pot.rho('Ni')], [pot.phi('NiNi')], [pot.embed('Ni')])
ranks = []
for r in (world - 1, 0, 1):
    lay = SlabLayout(lx, world, r, bench.RC, skin)
    ranks.append(SlabRank(model, lay, pos[lay.owned_mask(pos[:, 0])], ly, lz, 0, 'cuda'))
left, me, right = ranks
left.set_halo_counts(0, 0)
right.set_halo_counts(0, 0)
me.set_halo_counts(len(left.idx_r), len(right.idx_l))
me.recv_pos_l.copy_(left.pack_positions()[1])
me.recv_pos_r.copy_(right.pack_positions()[0])
for _ in range(3):
    me.build()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
ev0.record()
for _ in range(5):
    me.build()
ev1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 5 * 1e3
print(f"world {world} skin {skin}: owned {me.n_owned} halo {me.n_from_l + me.n_from_r} "
      f"build {ev0.elapsed_time(ev1) / 5:.3f} ms (device), {wall:.3f} ms (host wall), "
      f"nij {me.nbr.sizes()[0]}")
