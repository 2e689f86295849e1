"""One batched E+F+stress evaluation of 64 x 128-atom Be structures (AtomicNN G2+G4), the
command `ncu --set full -k regex:k_sf_` profiles (dev tool): python tools/sf_profile_run.py [medium]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_configs as bc     # noqa: E402
from tensoralloy_b200.precision import precision_scope   # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
with precision_scope(sys.argv[1] if len(sys.argv) > 1 else 'high'):
    print(bc.c2(64, True))
