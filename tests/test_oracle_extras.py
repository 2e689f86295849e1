"""CPU checks of the oracle pieces that have no golden in the reference (they only build
graphs there): the EAM training-step gradients against central finite differences of the
oracle loss, and the temperature-dependent heads' algebra / variable names."""
import numpy as np
import torch

from oracle import atomic as oat
from oracle import potentials as opot
from oracle import training as otr
from tensoralloy_b200.atoms import bulk_fcc
from tensoralloy_b200.nn.atomic import BeNN, SymmetryFunction, TemperatureDependentAtomicNN
from tensoralloy_b200.transformer import UniversalTransformer


def _structures():
    rng = np.random.default_rng(2)
    base = bulk_fcc('Ni', 3.55, (2, 2, 2))
    sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
    pos = base.positions + rng.normal(scale=0.08, size=base.positions.shape)
    return [dict(symbols=sym, positions=pos, cell=base.cell, pbc=[1, 1, 1],
                 energy=-4.2 * len(base), forces=rng.normal(scale=0.2, size=pos.shape),
                 stress=rng.normal(scale=0.01, size=6))]


def test_eam_training_gradients_match_finite_differences():
    structs = _structures()

    def run(shift=None):
        zj = opot.get_potential('zjw04')
        leaves = zj.track_parameters()
        if shift:
            (sec, key), dv = shift
            zj.params = {k: dict(v) for k, v in zj.params.items()}
            zj.params[sec][key] = zj.params[sec][key] + dv
        fns = {'rho': zj.rho, 'phi': zj.phi, 'embed': zj.embed}
        return otr.eam_loss_and_grads('alloy', ['Mo', 'Ni'], structs, fns,
                                      lambda: {f"{s}/{k}": t for (s, k), t in leaves.items()},
                                      5.0)
    loss, parts, grads = run()
    assert np.isfinite(loss) and parts['forces'] > 0
    for sec, key in (('Ni', 'A'), ('Mo', 'beta'), ('Ni', 'r_eq'), ('Mo', 'F1')):
        h = 1e-6
        lp = run(((sec, key), +h))[0]
        lm = run(((sec, key), -h))[0]
        fd = (lp - lm) / (2 * h)
        g = float(grads[f"{sec}/{key}"])
        assert abs(fd - g) < 2e-5 * max(1.0, abs(g)), (sec, key, fd, g)


def test_td_heads_algebra_and_variable_names():
    nn = TemperatureDependentAtomicNN(
        ['Be'], SymmetryFunction(['Be']), hidden_sizes=[8, 8],
        finite_temperature=dict(activation='tanh', layers=[12, 6], algo='Sommerfeld'))
    nn.attach_transformer(UniversalTransformer(['Be'], rcut=5.0, angular=True))
    nn.initialize_variables(seed=1)
    names = set(nn.variables)
    for head, layers in (('H', 1), ('S', 2), ('U', 2)):
        for k in range(layers):
            assert f"TD/Be/{head}/Conv1d{k + 1}/kernel" in names
        assert f"TD/Be/{head}/Output/kernel" in names
    assert nn.variables['TD/Be/H/Output/kernel'].shape[-1] == 6
    assert nn.variables['TD/Be/S/Conv1d1/kernel'].shape[-2] == 7     # [H, T]
    assert nn.is_finite_temperature and nn.variational_energy == 'free_energy'
    assert nn.as_dict()['finite_temperature']['algo'] == 'Sommerfeld'
    x = torch.rand(5, nn._dim(), dtype=torch.float64)
    p = nn.td_params('Be')
    U, S, F = oat.td_heads(x, 0.3, p, torch.float64)
    assert torch.allclose(F, U - 0.3 * S)
    p0 = dict(p, algo='default')
    S0 = oat.td_heads(x, 0.3, p0, torch.float64)[1]
    assert torch.allclose(S, 0.3 * S0)                                # Sommerfeld: S = T * net
    be = BeNN(['Be'], SymmetryFunction(['Be']), hidden_sizes=[8])
    be.attach_transformer(UniversalTransformer(['Be'], rcut=5.0, angular=True))
    be.initialize_variables(seed=1)
    assert 'TD/Be/S/Output/bias' not in be.variables               # beryllium.py:66
    assert 'TD/Be/U/Output/bias' in be.variables


def test_trainable_torch_forms_match_the_oracle_functions():
    """The torch forms the EAM trainer differentiates (nn/eam/training.py, product code)
    against the oracle's restatement of the same reference functions, on the CPU: zjw04
    (incl. Al-Cu mixing and the three embedding branches), sutton90, AgrawalBe, grimes."""
    from tensoralloy_b200.nn.eam import EamAlloyNN
    from tensoralloy_b200.nn.eam.training import _Functions
    r = torch.linspace(1.6, 6.4, 400, dtype=torch.float64)
    cases = [('zjw04', ['Al', 'Cu'], 'zjw04', torch.linspace(0.5, 60.0, 400, dtype=torch.float64)),
             ('sutton90', ['Ag'], 'sutton90', torch.linspace(0.5, 30.0, 200, dtype=torch.float64)),
             ('Be/1', ['Be'], 'Be/1', torch.linspace(0.05, 3.0, 200, dtype=torch.float64)),
             ('grimes', ['Pu'], 'grimes', torch.linspace(0.5, 30.0, 200, dtype=torch.float64))]
    rho60 = torch.linspace(0.5, 60.0, 400, dtype=torch.float64)
    cases += [('zjw04xc', ['Al', 'Cu'], 'zjw04xc', rho60), ('zjw04uxc', ['Be', 'Ni'], 'zjw04uxc', rho60),
              ('zjw04xcp', ['Mo', 'Ni'], 'zjw04xcp', rho60)]
    for name, els, oname, rho in cases:
        nn = EamAlloyNN(els, custom_potentials=name)
        fns = _Functions(nn, torch.float64, 'cpu', freeze_reference_fixed=False)
        ref = opot.get_potential(oname)
        for el in els:
            y = fns.get('rho', el)(r)
            assert torch.allclose(y, ref.rho(r, el), rtol=1e-13, atol=1e-15), (name, el, 'rho')
            y = fns.get('embed', el)(rho)
            assert torch.allclose(y, ref.embed(rho, el), rtol=1e-13, atol=1e-14), (name, el)
        for term in nn.unique_kbody_terms:
            y = fns.get('phi', term)(r)
            assert torch.allclose(y, ref.phi(r, term), rtol=1e-12, atol=1e-14), (name, term)
        # every variable is a leaf known by its reference name, trainable by default
        assert fns.params() and all(k.startswith('EAM/Shared/') for k in fns.named)
    # fixed_functions freeze the variables of that function only
    nn = EamAlloyNN(['Ag'], custom_potentials='sutton90', fixed_functions=['Ag.rho'])
    fns = _Functions(nn, torch.float64, 'cpu')
    fns.get('rho', 'Ag'), fns.get('phi', 'AgAg')
    assert not fns.named['EAM/Shared/Ag/a'].requires_grad
    assert fns.named['EAM/Shared/AgAg/b'].requires_grad


def test_trainable_mishinh_and_constant_msah11_forms():
    """MishinH embed / dipole / quadrupole (mishin.py:20-315) as trainable torch forms and the
    constant Mendelev Al-Fe functions (msah11.py:28-424) against the oracle restatement."""
    from tensoralloy_b200.nn.eam import AdpNN, EamFsNN
    from tensoralloy_b200.nn.eam.training import _Functions
    r = torch.linspace(1.2, 6.6, 500, dtype=torch.float64)
    rho = torch.linspace(0.0, 3.0, 300, dtype=torch.float64)
    nn = AdpNN(['Mo', 'Ni'], custom_potentials={
        'Mo': {'rho': 'zjw04', 'embed': 'mishinh'}, 'Ni': {'rho': 'zjw04', 'embed': 'zjw04'},
        'MoMo': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'},
        'MoNi': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'},
        'NiNi': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'}})
    fns = _Functions(nn, torch.float64, 'cpu')
    ref = opot.get_potential('mishinh')
    y = fns.get('embed', 'Mo')(rho)
    assert torch.allclose(y, ref.embed(rho, 'Mo'), rtol=1e-13, atol=1e-15)
    for term in ('MoMo', 'MoNi', 'NiNi'):
        assert torch.allclose(fns.get('dipole', term)(r), ref.dipole(r, term),
                              rtol=1e-13, atol=1e-16), term
        assert torch.allclose(fns.get('quadrupole', term)(r), ref.quadrupole(r, term),
                              rtol=1e-13, atol=1e-16), term
    assert fns.named['ADP/Shared/Mo/s5'].requires_grad
    assert fns.named['ADP/Shared/NiNi/q2'].requires_grad
    # gradients flow to the shared variables
    fns.get('dipole', 'NiNi')(r).sum().backward()
    assert fns.named['ADP/Shared/NiNi/d1'].grad is not None
    # the reference never trains the embedding parameters / r_eq of zjw04 (zjw04.py:174-177)
    fns.get('embed', 'Ni'), fns.get('rho', 'Ni')
    assert not fns.named['ADP/Shared/Ni/F0'].requires_grad
    assert not fns.named['ADP/Shared/Ni/r_eq'].requires_grad
    assert fns.named['ADP/Shared/Ni/f_eq'].requires_grad

    fs = EamFsNN(['Al', 'Fe'], custom_potentials='msah11')
    fns = _Functions(fs, torch.float64, 'cpu')
    ref = opot.get_potential('msah11')
    rr = torch.linspace(0.5, 6.6, 700, dtype=torch.float64)
    for term in ('AlAl', 'AlFe', 'FeFe'):
        assert torch.allclose(fns.get('phi', term)(rr), ref.phi(rr, term),
                              rtol=1e-13, atol=1e-14), term
    for term in ('AlAl', 'AlFe', 'FeAl', 'FeFe'):
        assert torch.allclose(fns.get('rho', term)(rr), ref.rho(rr, term),
                              rtol=1e-13, atol=1e-16), term
    rho = torch.linspace(0.0, 60.0, 300, dtype=torch.float64)
    for el in ('Al', 'Fe'):
        assert torch.allclose(fns.get('embed', el)(rho), ref.embed(rho, el),
                              rtol=1e-13, atol=1e-15), el
    assert not fns.params()              # constants only


def test_grap_T_and_M_full_and_unique_forms_agree():
    """nn/atomic/tests/test_grap.py:152-200 (`test_T_and_M`): sum_d T_dm M_d is the same
    for the full 3^m tensors and for the unique tuples with multiplicities, m = 0..3."""
    import torch
    from oracle import atomic as oat
    g = torch.Generator().manual_seed(5)
    D = torch.randn(40, 3, generator=g, dtype=torch.float64)
    u = D / D.norm(dim=1, keepdim=True)
    for mm in range(4):
        V = torch.as_tensor(oat.grap_T_dm_full(mm)).T @ oat.grap_moment_tensor_full(u, mm)
        v = torch.as_tensor(oat.grap_multiplicity_tensor(mm)).T @ oat.grap_moment_coeff(u, mm)
        assert V.shape == v.shape == (mm + 1, 40)
        assert float((V - v).abs().max()) < 1e-8          # the reference's delta
        assert float((V - v).abs().max()) < 1e-14


def test_grap_new_mode_equals_legacy_mode():
    """nn/atomic/tests/test_grap.py:46-105 (Be/W, pexp, polynomial cutoff, moments 0-2,
    delta 1e-6) and :108-149 (`test_sign`: Fe bcc lattice sweep, morse, moments 0-1,
    delta 1e-8): the two formulations of the reference give the same descriptors."""
    import torch
    from oracle import atomic as oat
    from oracle import neighbor as onl
    rng = np.random.default_rng(12)
    # Be hcp 2x2x2 with eight W atoms, rattled
    a, c = 2.29, 3.58
    cell0 = np.array([[a, 0, 0], [-a / 2, a * np.sqrt(3) / 2, 0], [0, 0, c]])
    base = np.array([[0, 0, 0], [1 / 3, 2 / 3, 1 / 2]]) @ cell0
    pos = np.array([b + np.array([i, j, k]) @ cell0 for i in range(2) for j in range(2)
                    for k in range(2) for b in base]) + rng.random((16, 3)) * 0.1
    cell = cell0 * 2
    types = np.array([0] * 8 + [1] * 8)
    rc = 5.0
    i, j, S = onl.neighbor_list(pos, cell, [1, 1, 1], rc)[:3]
    R, h = torch.as_tensor(pos), torch.as_tensor(cell)
    grid = [(rl, pl) for rl, pl in zip(np.linspace(1.0, 4.0, 4), [1.0, 2.0, 3.0, 2.5])]
    legacy = oat.grap_descriptors(['Be', 'W'], types, R, h, i, j, S, rc, 'pexp', grid,
                                  (0, 1, 2), cutoff='polynomial')
    new = oat.grap_descriptors_new_mode(['Be', 'W'], types, R, h, i, j, S, rc, 'pexp', grid,
                                        2, cutoff='polynomial')
    assert legacy.shape == new.shape == (16, 2 * 4 * 3)
    assert float((legacy - new).abs().max()) < 1e-10
    # Fe bcc sweep, morse (sums change sign along the sweep -> exercises sign(P))
    mgrid = [(1.0, 1.0, r0) for r0 in (3.3, 3.4, 3.5)]
    seen_negative = False
    for x in range(-20, 21, 4):
        lat = 2.87 * (1.0 + x / 100.0)
        pos = np.array([[0, 0, 0], [0.5, 0.5, 0.5]]) * lat
        cell = np.eye(3) * lat
        i, j, S = onl.neighbor_list(pos, cell, [1, 1, 1], 6.0)[:3]
        R, h = torch.as_tensor(pos), torch.as_tensor(cell)
        y = oat.grap_descriptors(['Fe'], np.zeros(2, int), R, h, i, j, S, 6.0, 'morse',
                                 mgrid, (0, 1))
        z = oat.grap_descriptors_new_mode(['Fe'], np.zeros(2, int), R, h, i, j, S, 6.0,
                                          'morse', mgrid, 1)
        assert y.shape == z.shape == (2, 6)
        assert float((y - z).abs().max()) < 1e-8
        seen_negative |= bool((y[:, 0::2] < 0).any())
    assert seen_negative


def test_grap_new_mode_symmetric_is_traceless_form():
    """grap.py:485-494: symmetric T_dm subtracts P_0^2 / 3 from the m = 2 entry and
    3/5 sum_a P_a^2 from the m = 3 entry (the traces of the moment tensors)."""
    import torch
    from oracle import atomic as oat
    from oracle import neighbor as onl
    rng = np.random.default_rng(3)
    pos = rng.random((6, 3)) * 4.0
    cell = np.eye(3) * 4.0
    i, j, S = onl.neighbor_list(pos, cell, [1, 1, 1], 3.5)[:3]
    R, h = torch.as_tensor(pos), torch.as_tensor(cell)
    args = (['Be'], np.zeros(6, int), R, h, i, j, S, 3.5, 'sf', [(2.0, 0.0), (8.0, 1.0)], 3)
    plain = oat.grap_descriptors_new_mode(*args).reshape(6, 2, 4)
    sym = oat.grap_descriptors_new_mode(*args, symmetric=True).reshape(6, 2, 4)
    assert torch.allclose(sym[..., :2], plain[..., :2], atol=0, rtol=0)
    assert torch.allclose(sym[..., 2], plain[..., 2] - plain[..., 0] ** 2 / 3.0, atol=1e-13)
    assert torch.allclose(sym[..., 3], plain[..., 3] - 0.6 * plain[..., 1], atol=1e-13)


def test_grap_new_mode_oracle_forces_and_stress_by_finite_differences():
    """The new-mode restatement (moment 3, traceless form, and the `nn` filter network) that
    anchors the GPU parity tests: autograd forces = -dE/dR and stress = (dE/d strain) / V by
    central differences."""
    import torch
    from oracle import atomic as oat
    from tensoralloy_b200.atoms import bulk_fcc
    rng = np.random.default_rng(2)
    base = bulk_fcc('Ni', 3.6, (2, 2, 2))
    pos = base.positions + rng.normal(scale=0.08, size=base.positions.shape)
    cell = np.asarray(base.cell)
    sym = ['Mo' if k % 3 == 0 else 'Ni' for k in range(len(pos))]
    elements, rc = ['Mo', 'Ni'], 4.5
    K = 3

    def net(n_in, sizes, seed):
        r = np.random.default_rng(seed)
        dims = [n_in] + sizes
        W = [r.normal(size=(dims[k], dims[k + 1])) * 0.4 for k in range(len(sizes))]
        b = [r.normal(size=dims[k + 1]) * 0.1 for k in range(len(sizes) - 1)] + [None]
        return W, b

    fW, fb = net(1, [6, 6, K], 1)
    cases = [dict(algorithm='pexp', grid=[(1.5, 2.0), (2.5, 3.0), (2.0, 1.0)], symmetric=True),
             dict(algorithm='nn', grid=dict(weights=fW, biases=fb, activation='tanh',
                                            use_resnet_dt=True), symmetric=False)]
    for case in cases:
        grap = dict(moments=[0, 1, 2, 3], cutoff='cosine', new_mode=True, **case)
        dim = 2 * K * 4
        params = {}
        for k, el in enumerate(elements):
            W, b = net(dim, [8, 1], 10 + k)
            params[el] = dict(weights=[w * 0.2 for w in W], biases=b, activation='softplus',
                              use_resnet_dt=False, out_bias=None)

        def run(p, c):
            return oat.atomic_evaluate(elements, sym, p, c, [1, 1, 1], rc, params,
                                       angular=False, grap=grap)
        ref = run(pos, cell)
        assert np.abs(ref['forces']).max() > 1e-3
        h = 1e-5
        for a, c in ((0, 0), (7, 1), (20, 2)):
            p1, p2 = pos.copy(), pos.copy()
            p1[a, c] += h
            p2[a, c] -= h
            fd = -(run(p1, cell)['energy'] - run(p2, cell)['energy']) / (2 * h)
            assert abs(fd - ref['forces'][a, c]) < 1e-7 * max(1.0, abs(fd)), (case['algorithm'], a)
        eps, vol = 1e-6, abs(np.linalg.det(cell))
        for (i, j), v in (((0, 0), 0), ((1, 2), 3)):
            st = np.zeros((3, 3))
            st[i, j] = st[j, i] = eps if i == j else eps / 2
            Fp, Fm = np.eye(3) + st, np.eye(3) - st
            de = (run(pos @ Fp, cell @ Fp)['energy'] - run(pos @ Fm, cell @ Fm)['energy']) \
                / (2 * eps)
            assert abs(de / vol - ref['stress'][v]) < 1e-7 * max(1.0, abs(de / vol)), \
                (case['algorithm'], v)
