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
