"""Kernel-by-kernel time of one neighbour-list build of the 1 M-atom headline (dev tool; run it
under `ncu --metrics gpu__time_duration.sum`): python tools/build_breakdown.py [skin]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch          # noqa: E402
import bench          # noqa: E402
from tensoralloy_b200 import _lib   # noqa: E402

skin = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
pos, cell = bench.make_lattice(63)
d_pos = torch.from_numpy(pos).cuda()
nbr = _lib.NeighborList()
nbr.set_skin(skin)
for _ in range(3):
    nbr.build(d_pos, None, cell, [1, 1, 1], bench.RC)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(5):
    nbr.build(d_pos, None, cell, [1, 1, 1], bench.RC)
ev1.record()
torch.cuda.synchronize()
print(f"skin {skin}: build {ev0.elapsed_time(ev1) / 5:.3f} ms, nij {nbr.sizes()[0]}")
