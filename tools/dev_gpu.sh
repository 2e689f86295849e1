mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_atomic_gpu.py -m gpu -x -q -s -k "tensor_cores" 2>&1 | grep "dE/N\|passed\|failed\|Error" | tail -12
python - <<'PY'
# throughput of the MLP stage alone: tensor cores vs warp-per-atom, 'medium', large system
import os, time, numpy as np, torch
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.nn.atomic import AtomicNN, SymmetryFunction
from tensoralloy_b200.precision import precision_scope, get_float_dtype
from tensoralloy_b200.transformer import UniversalTransformer
base = bulk_fcc('Ni', 3.6, (30, 30, 30))          # 108 000 atoms
rng = np.random.default_rng(5)
atoms = Atoms(['Ni'] * len(base), base.positions + rng.normal(scale=0.05, size=base.positions.shape), base.cell, True)
with precision_scope('medium'):
    nn = AtomicNN(['Ni'], SymmetryFunction(['Ni']), minmax_scale=False, export_properties=('energy', 'forces', 'stress'))
    clf = UniversalTransformer(['Ni'], rcut=4.6, acut=4.0, angular=True)
    nn.attach_transformer(clf); nn.initialize_variables()
    feats = clf.get_constant_features(atoms)
    model = nn._device_model(); n = len(atoms)
    e = torch.zeros(16, dtype=torch.float64, device='cuda'); f = torch.zeros((n, 3), dtype=torch.float64, device='cuda')
    for tc in ('0', '1'):
        os.environ['TAB_MLP_TC'] = tc
        for _ in range(3): model.eval(feats.nbr, 1, energy=e[0:1], forces=f, virial=e[1:10])
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): model.eval(feats.nbr, 1, energy=e[0:1], forces=f, virial=e[1:10])
        torch.cuda.synchronize(); print('TAB_MLP_TC=' + tc, 'eval ms', 1e2 * (time.perf_counter() - t0), 'E', e[0].item())
PY
TAB_MLP_TC=1 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_mlp" -c 4 --csv --log-file gpurun_out/mlp_tc_ncu.csv python - <<'PY' > /dev/null 2>&1
import os, numpy as np, torch
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.nn.atomic import AtomicNN, SymmetryFunction
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer
base = bulk_fcc('Ni', 3.6, (30, 30, 30))
rng = np.random.default_rng(5)
atoms = Atoms(['Ni'] * len(base), base.positions + rng.normal(scale=0.05, size=base.positions.shape), base.cell, True)
with precision_scope('medium'):
    nn = AtomicNN(['Ni'], SymmetryFunction(['Ni']), minmax_scale=False, export_properties=('energy', 'forces', 'stress'))
    clf = UniversalTransformer(['Ni'], rcut=4.6, acut=4.0, angular=True)
    nn.attach_transformer(clf); nn.initialize_variables()
    feats = clf.get_constant_features(atoms)
    model = nn._device_model(); n = len(atoms)
    e = torch.zeros(16, dtype=torch.float64, device='cuda'); f = torch.zeros((n, 3), dtype=torch.float64, device='cuda')
    for tc in ('1', '0'):
        os.environ['TAB_MLP_TC'] = tc
        for _ in range(2): model.eval(feats.nbr, 1, energy=e[0:1], forces=f, virial=e[1:10])
    torch.cuda.synchronize()
PY
cat gpurun_out/mlp_tc_ncu.csv | grep -v "^==" | cut -d, -f5,11- | head -20
