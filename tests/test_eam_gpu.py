"""EAM E / F / virial on the GPU vs the oracle (torch float64 autograd).
Tolerances from BASELINE.json north_star: 1e-10 eV/atom, 1e-8 eV/A (float64);
1e-5 relative (float32)."""
import numpy as np
import pytest
import torch

from oracle import eam as oeam
from oracle import potentials as opot
from tensoralloy_b200 import _lib
from tensoralloy_b200.atoms import bulk_fcc
from tensoralloy_b200.nn.eam.potentials import get_potential

pytestmark = pytest.mark.gpu


def gpu_eam(pot_name, elements, symbols, pos, cell, pbc, rc, precision=0,
            kind=_lib.EAM_ALLOY):
    elements = sorted(elements)
    pot = get_potential(pot_name)
    n_el = len(elements)
    rho = [pot.rho(f'{a}{b}') for a in elements for b in elements]
    phi = [pot.phi(''.join(sorted([a, b]))) for a in elements for b in elements]
    embed = [pot.embed(a) for a in elements]
    model = _lib.EamModel(kind, n_el, rho, phi, embed)
    nl = _lib.NeighborList()
    n = len(pos)
    d_pos = torch.tensor(pos, dtype=torch.float64, device='cuda').contiguous()
    d_t = torch.tensor([elements.index(s) for s in symbols], dtype=torch.int32,
                       device='cuda')
    nl.build(d_pos, d_t, cell, pbc, rc)
    e = torch.zeros(1, dtype=torch.float64, device='cuda')
    ea = torch.zeros(n, dtype=torch.float64, device='cuda')
    f = torch.zeros((n, 3), dtype=torch.float64, device='cuda')
    v = torch.zeros(9, dtype=torch.float64, device='cuda')
    model.eval(nl, precision, e, ea, f, v)
    torch.cuda.synchronize()
    return (e.cpu().numpy()[0], ea.cpu().numpy(), f.cpu().numpy(),
            v.cpu().numpy().reshape(3, 3))


def compare(pot_name, elements, symbols, pos, cell, pbc, rc):
    ref = oeam.eam_evaluate(opot.get_potential(pot_name), 'alloy', elements,
                            symbols, pos, cell, pbc, rc)
    e, ea, f, v = gpu_eam(pot_name, elements, symbols, pos, cell, pbc, rc)
    n = len(pos)
    assert abs(e - ref['energy']) / n < 1e-10
    assert np.abs(ea - ref['energy/atom']).max() < 1e-10
    assert np.abs(f - ref['forces']).max() < 1e-8
    assert np.abs(v - ref['virial']).max() / n < 1e-8
    # float32 'medium'
    e32, _, f32, v32 = gpu_eam(pot_name, elements, symbols, pos, cell, pbc, rc,
                               precision=1)
    assert abs(e32 - ref['energy']) <= 1e-5 * abs(ref['energy'])
    fscale = max(np.abs(ref['forces']).max(), 1e-3)
    print('float32: dE/|E|', abs(e32 - ref['energy']) / abs(ref['energy']),
          'dF/Fmax', np.abs(f32 - ref['forces']).max() / fscale, 'Fmax', fscale)
    # 1e-5 of the force scale (measured <= 2.5e-6); on the perfect lattice the forces vanish
    # by symmetry and the scale is that of a rattled one (0.05 eV/A)
    assert np.abs(f32 - ref['forces']).max() <= 1e-5 * max(fscale, 0.05)
    return ref


def test_ni_fcc_256():
    atoms = bulk_fcc('Ni', 3.52, (4, 4, 4))
    sym = atoms.get_chemical_symbols()
    ref = compare('zjw04', ['Ni'], sym, atoms.positions, atoms.cell, [1, 1, 1], 6.5)
    # fcc E/atom = -4.44999667 eV at rc 6.5
    assert abs(ref['energy'] / 256 + 4.44999667) < 1e-7
    rng = np.random.default_rng(611)
    pos = atoms.positions + rng.normal(scale=0.05, size=atoms.positions.shape)
    compare('zjw04', ['Ni'], sym, pos, atoms.cell, [1, 1, 1], 6.5)
    compare('zjw04', ['Ni'], sym, pos, atoms.cell, [1, 1, 1], 6.0)
    compare('zjw04xc', ['Ni'], sym, pos, atoms.cell, [1, 1, 1], 6.5)


def test_mo_ni_alloy():
    atoms = bulk_fcc('Ni', 3.6, (3, 3, 3))
    rng = np.random.default_rng(7)
    sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(atoms))]
    pos = atoms.positions + rng.normal(scale=0.08, size=atoms.positions.shape)
    compare('zjw04', ['Mo', 'Ni'], sym, pos, atoms.cell, [1, 1, 1], 6.0)
    compare('zjw04xcp', ['Mo', 'Ni'], sym, pos, atoms.cell, [1, 1, 1], 6.0)


def test_small_cell_images():
    atoms = bulk_fcc('Ni', 3.52, (2, 2, 2))
    rng = np.random.default_rng(1)
    pos = atoms.positions + rng.normal(scale=0.03, size=atoms.positions.shape)
    compare('zjw04', ['Ni'], atoms.get_chemical_symbols(), pos, atoms.cell,
            [1, 1, 1], 6.5)


# ---------------------------------------------------------------------------
# further potentials, Finnis-Sinclair and ADP through the model classes
# ---------------------------------------------------------------------------
def _calc_compare(nn, atoms, oracle_pot, kind, rc, fns=None, tol_e=1e-10,
                  tol_f=1e-8):
    from tensoralloy_b200.calculator import TensorAlloyCalculator
    from tensoralloy_b200.precision import precision_scope
    from tensoralloy_b200.transformer import UniversalTransformer
    with precision_scope('high'):
        nn.attach_transformer(UniversalTransformer(nn.elements, rcut=rc))
        calc = TensorAlloyCalculator(nn)
        calc.calculate(atoms, properties=['energy', 'forces', 'stress'])
        e = calc.results['energy']
        f = calc.get_forces(atoms)
        s = calc.get_stress(atoms)
        ea = calc.get_atomic(atoms)
    ref = oeam.eam_evaluate(oracle_pot, kind, nn.elements,
                            atoms.get_chemical_symbols(), atoms.positions,
                            atoms.cell, atoms.pbc, rc, fns=fns)
    n = len(atoms)
    assert abs(e - ref['energy']) / n < tol_e
    assert np.abs(ea - ref['energy/atom']).max() < tol_e * 10
    assert np.abs(f - ref['forces']).max() < tol_f
    assert np.abs(s - ref['stress']).max() < tol_f
    return ref


def _rattled(symbol, a, rep, seed, sigma=0.06):
    atoms = bulk_fcc(symbol, a, rep)
    rng = np.random.default_rng(seed)
    atoms.positions += rng.normal(scale=sigma, size=atoms.positions.shape)
    return atoms


def test_sutton90_agrawal_grimes():
    from tensoralloy_b200.nn.eam import EamAlloyNN
    _calc_compare(EamAlloyNN(['Ag'], custom_potentials='sutton90'),
                  _rattled('Ag', 4.09, (3, 3, 3), 1), opot.get_potential('sutton90'),
                  'alloy', 6.0)
    _calc_compare(EamAlloyNN(['Be'], custom_potentials='Be/1'),
                  _rattled('Be', 3.2, (3, 3, 3), 2), opot.get_potential('Be/1'),
                  'alloy', 5.0)
    _calc_compare(EamAlloyNN(['Pu'], custom_potentials='grimes'),
                  _rattled('Pu', 4.6, (3, 3, 3), 3), opot.get_potential('grimes'),
                  'alloy', 6.0, tol_f=1e-7)


def test_finnis_sinclair_ordered_pair_density():
    """EAM/FS: rho depends on the ordered (centre, neighbour) pair (fs.py:180-203).
    Built at the table level with four distinct density functions."""
    pz = get_potential('zjw04')
    els = ['Mo', 'Ni']
    src = {'MoMo': 'Mo', 'MoNi': 'Cu', 'NiMo': 'Al', 'NiNi': 'Ni'}
    rho = [pz.rho(src[a + b]) for a in els for b in els]
    phi = [pz.phi(''.join(sorted([a, b]))) for a in els for b in els]
    embed = [pz.embed(a) for a in els]
    model = _lib.EamModel(_lib.EAM_FS, 2, rho, phi, embed)
    atoms = _rattled('Ni', 3.6, (3, 3, 3), 5)
    rng = np.random.default_rng(9)
    sym = ['Mo' if x < 0.5 else 'Ni' for x in rng.random(len(atoms))]
    n = len(atoms)
    nl = _lib.NeighborList()
    d_pos = torch.tensor(atoms.positions, device='cuda')
    d_t = torch.tensor([els.index(s) for s in sym], dtype=torch.int32, device='cuda')
    nl.build(d_pos, d_t, atoms.cell, [1, 1, 1], 6.0)
    e = torch.zeros(1, dtype=torch.float64, device='cuda')
    f = torch.zeros((n, 3), dtype=torch.float64, device='cuda')
    v = torch.zeros(9, dtype=torch.float64, device='cuda')
    model.eval(nl, 0, energy=e, forces=f, virial=v)
    op = opot.get_potential('zjw04')
    fns = {'rho': lambda r, term: op.rho(r, src[term])}
    ref = oeam.eam_evaluate(op, 'fs', els, sym, atoms.positions, atoms.cell,
                            [1, 1, 1], 6.0, fns=fns)
    assert abs(e.item() - ref['energy']) / n < 1e-10
    assert np.abs(f.cpu().numpy() - ref['forces']).max() < 1e-8
    assert np.abs(v.cpu().numpy().reshape(3, 3) - ref['virial']).max() / n < 1e-8


def test_adp_dipole_quadrupole():
    """AdpNN: zjw04 rho/phi/embed + MishinH dipole/quadrupole (adp.py:315-586),
    one and two species (per-term squaring of the moments)."""
    from tensoralloy_b200.nn.eam import AdpNN
    om = opot.get_potential('mishinh')
    oz = opot.get_potential('zjw04')
    fns = {'rho': oz.rho, 'phi': oz.phi, 'embed': oz.embed,
           'dipole': om.dipole, 'quadrupole': om.quadrupole}
    cp = {'Ni': {'rho': 'zjw04', 'embed': 'zjw04'},
          'NiNi': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'}}
    nn = AdpNN(['Ni'], custom_potentials=cp)
    _calc_compare(nn, _rattled('Ni', 3.52, (3, 3, 3), 11), None, 'adp', 6.0, fns=fns)
    cp2 = {'Mo': {'rho': 'zjw04', 'embed': 'zjw04'},
           'Ni': {'rho': 'zjw04', 'embed': 'zjw04'},
           'MoMo': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'},
           'MoNi': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'},
           'NiNi': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'}}
    atoms = _rattled('Ni', 3.6, (3, 3, 3), 12)
    rng = np.random.default_rng(13)
    from tensoralloy_b200.atoms import Atoms
    sym = ['Mo' if x < 0.45 else 'Ni' for x in rng.random(len(atoms))]
    atoms = Atoms(sym, atoms.positions, atoms.cell, True)
    _calc_compare(AdpNN(['Mo', 'Ni'], custom_potentials=cp2), atoms, None, 'adp', 6.0,
                  fns=fns)


def test_elastic_constants_known_answers():
    """zjw04 Ni fcc, rc 6.0: C11 = 247, C12 = 147, C44 = 125 GPa (+-1), the values
    asserted by the reference in nn/constraint/tests/test_elastic.py:49-56
    (246.61 / 147.15 / 124.72 in tests/test_calculator.py:101-108).  Obtained here
    by central finite differences of the GPU stress under homogeneous strain:
    pins the sign and units of the virial / stress."""
    from tensoralloy_b200.atoms import GPa
    from tensoralloy_b200.calculator import TensorAlloyCalculator
    from tensoralloy_b200.nn.eam import EamAlloyNN
    from tensoralloy_b200.precision import precision_scope
    from tensoralloy_b200.transformer import UniversalTransformer
    with precision_scope('high'):
        nn = EamAlloyNN(['Ni'], custom_potentials='zjw04',
                        export_properties=['energy', 'forces', 'stress'])
        nn.attach_transformer(UniversalTransformer(['Ni'], rcut=6.0))
        calc = TensorAlloyCalculator(nn)
        base = bulk_fcc('Ni', 3.52, (3, 3, 3))

        def stress(eps):
            a = base.copy()
            F = np.eye(3) + eps
            a.set_cell(base.cell @ F, scale_atoms=True)
            return calc.get_stress(a, voigt=False)

        h = 1e-4
        e = np.zeros((3, 3)); e[0, 0] = h
        d = (stress(e) - stress(-e)) / (2 * h) / GPa
        c11, c12 = d[0, 0], d[1, 1]
        e = np.zeros((3, 3)); e[1, 2] = e[2, 1] = h / 2
        d = (stress(e) - stress(-e)) / (2 * h) / GPa
        c44 = d[1, 2]
    assert abs(c11 - 246.61) < 0.5 and abs(c12 - 147.15) < 0.5 and abs(c44 - 124.72) < 0.5


# --- lists from the block-per-tile build (large systems; forced here on small ones) ---
def test_tile_lists_feed_the_eam_passes(monkeypatch):
    atoms = bulk_fcc('Ni', 3.52, (8, 7, 6))           # 1344 atoms, odd bin counts
    rng = np.random.default_rng(611)
    pos = atoms.positions + rng.normal(scale=0.05, size=atoms.positions.shape)
    sym = atoms.get_chemical_symbols()
    monkeypatch.setenv('TAB_NBR_MODE', 'tile')
    compare('zjw04', ['Ni'], sym, pos, atoms.cell, [1, 1, 1], 6.5)
    # non-periodic slab (vacuum along z) and a two-species system (regrouped rows)
    cell = atoms.cell.copy()
    cell[2, 2] += 12.0
    compare('zjw04', ['Ni'], sym, pos, cell, [1, 1, 0], 6.5)
    sym2 = ['Mo' if x < 0.5 else 'Ni' for x in rng.random(len(pos))]
    compare('zjw04', ['Mo', 'Ni'], sym2, pos, atoms.cell, [1, 1, 1], 6.5)


def test_tile_lists_large_lattice(monkeypatch):
    # default dispatch at 32 000 atoms (tile build) vs the thread-per-atom build:
    # same rows in the same order -> identical results
    from tensoralloy_b200.atoms import fcc_positions
    pos, cell = fcc_positions(3.52, 20, 20, 20)
    pos = pos + np.random.default_rng(611).normal(scale=0.05, size=pos.shape)
    sym = ['Ni'] * len(pos)
    e1, ea1, f1, v1 = gpu_eam('zjw04', ['Ni'], sym, pos, cell, [1, 1, 1], 6.5)
    monkeypatch.setenv('TAB_NBR_MODE', 'thread')
    e0, ea0, f0, v0 = gpu_eam('zjw04', ['Ni'], sym, pos, cell, [1, 1, 1], 6.5)
    assert e1 == e0
    np.testing.assert_array_equal(f1, f0)
    np.testing.assert_array_equal(ea1, ea0)


# --- 'nn' functions (eam.py:174-190): MLP-parametrised rho / phi / embed / u / w ---------
def _nn_oracle_fns(nn, elements):
    """Oracle callables for the functions of `nn` that are 'nn' (torch MLP with the
    model's weights); the others stay with the named empirical potential."""
    from oracle import atomic as oat
    provider = nn._nn

    def make(fn, key):
        arrays = provider.weights(fn, key)
        W = [torch.tensor(a) for a in arrays[0:-1:2]] + [torch.tensor(arrays[-1])]
        b = [torch.tensor(a) for a in arrays[1:-1:2]]
        return lambda x: oat.mlp(x[:, None], W, b, nn._activation)

    fns = {}
    pots = nn.potentials
    zj = opot.get_potential('zjw04')

    def dispatch(fn, base):
        def call(x, key):
            if fn in ('rho', 'embed'):
                section = key[-2:] if (fn == 'rho' and nn.tag == 'fs') else key
                if fn == 'rho' and nn.tag != 'fs':
                    section = key
            else:
                section = key
            name = pots[section][fn]
            if name == 'nn':
                return make(fn, section)(x)
            return getattr(zj, base)(x, key)
        return call

    for fn in ('rho', 'phi', 'embed', 'dipole', 'quadrupole'):
        fns[fn] = dispatch(fn, fn)
    return fns


@pytest.mark.parametrize("case", ["all_nn", "mixed", "adp_nn"])
def test_nn_parametrised_functions(case):
    from tensoralloy_b200.atoms import Atoms
    from tensoralloy_b200.calculator import TensorAlloyCalculator
    from tensoralloy_b200.nn.eam import AdpNN, EamAlloyNN
    from tensoralloy_b200.precision import precision_scope
    from tensoralloy_b200.transformer import UniversalTransformer
    base = bulk_fcc('Ni', 3.6, (3, 3, 3))
    rng = np.random.default_rng(7)
    sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
    atoms = Atoms(sym, base.positions + rng.normal(scale=0.08, size=base.positions.shape),
                  base.cell, True)
    elements = ['Mo', 'Ni']
    with precision_scope('high'):
        if case == "all_nn":
            nn = EamAlloyNN(elements, hidden_sizes=[16, 8])          # every function 'nn'
        elif case == "mixed":
            nn = EamAlloyNN(elements, custom_potentials={
                'Ni': {'rho': 'zjw04', 'embed': 'nn'}, 'Mo': {'rho': 'nn', 'embed': 'zjw04'},
                'MoNi': {'phi': 'nn'}, 'NiNi': {'phi': 'zjw04'}, 'MoMo': {'phi': 'zjw04'}},
                hidden_sizes={'Ni': {'embed': [24]}, 'Mo': {'rho': [12, 12, 6]},
                              'MoNi': {'phi': [32, 16]}})
        else:
            nn = AdpNN(elements, custom_potentials={
                'Ni': {'rho': 'zjw04', 'embed': 'zjw04'}, 'Mo': {'rho': 'zjw04', 'embed': 'zjw04'},
                'MoNi': {'phi': 'zjw04', 'dipole': 'nn', 'quadrupole': 'nn'},
                'NiNi': {'phi': 'zjw04', 'dipole': 'nn', 'quadrupole': 'nn'},
                'MoMo': {'phi': 'zjw04', 'dipole': 'nn', 'quadrupole': 'nn'}},
                hidden_sizes=[8, 8])
        nn.attach_transformer(UniversalTransformer(elements, rcut=5.0))
        nn.initialize_variables(seed=3)
        for name, value in list(nn.variables.items()):
            if name.endswith('Output/kernel'):
                nn.set_variable(name, value * (0.02 if case == "adp_nn" else 0.2))
        calc = TensorAlloyCalculator(nn)
        calc.calculate(atoms, properties=['energy', 'forces', 'stress'])
        e, f, s = calc.results['energy'], calc.get_forces(atoms), calc.get_stress(atoms)
    kind = 'adp' if case == "adp_nn" else 'alloy'
    ref = oeam.eam_evaluate(opot.get_potential('zjw04'), kind, elements, sym,
                            atoms.positions, atoms.cell, [1, 1, 1], 5.0,
                            fns=_nn_oracle_fns(nn, elements))
    n = len(atoms)
    print(case, 'E/N', ref['energy'] / n, 'dE/N', abs(e - ref['energy']) / n, 'dF',
          np.abs(f - ref['forces']).max(), 'Fmax', np.abs(ref['forces']).max())
    assert abs(e - ref['energy']) / n < 1e-10
    assert np.abs(f - ref['forces']).max() < 1e-8
    assert np.abs(s - ref['stress']).max() < 1e-8
    assert np.abs(ref['forces']).max() > 1e-3
    with precision_scope('medium'):
        calc32 = TensorAlloyCalculator(nn)
        calc32.calculate(atoms, properties=['energy', 'forces'])
    assert abs(calc32.results['energy'] - ref['energy']) <= 2e-5 * max(abs(ref['energy']), 1.0)


def test_msah11_al_fe_finnis_sinclair():
    """AlFeMsah11 (nn/eam/potentials/msah11.py:28-424) through EamFsNN: piecewise pair
    function from the coefficient pool, truncated-power densities, both embeddings.  The
    oracle restatement is pinned to the reference's golden table (test_oracle_golden)."""
    from tensoralloy_b200.atoms import Atoms
    from tensoralloy_b200.calculator import TensorAlloyCalculator
    from tensoralloy_b200.nn.eam import EamFsNN
    from tensoralloy_b200.precision import precision_scope
    pot = opot.get_potential('msah11')
    rng = np.random.default_rng(17)
    for a0, frac, sigma in ((4.05, 0.25, 0.08), (3.7, 0.7, 0.15)):
        base = bulk_fcc('Al', a0, (3, 3, 3))
        sym = ['Fe' if x < frac else 'Al' for x in rng.random(len(base))]
        atoms = Atoms(sym, base.positions + rng.normal(scale=sigma, size=base.positions.shape),
                      base.cell, True)
        nn = EamFsNN(['Al', 'Fe'], custom_potentials='msah11',
                     export_properties=('energy', 'forces', 'stress'))
        ref = _calc_compare(nn, atoms, pot, 'fs', 6.5)
        assert np.abs(ref['forces']).max() > 0.1
        with precision_scope('medium'):
            c32 = TensorAlloyCalculator(nn)
            c32.calculate(atoms, properties=['energy', 'forces'])
            # the reference itself warns against float32 for this potential
            # (msah11.py:23-25): the polynomial tails cancel to ~1e-4 relative in float32
            assert abs(c32.results['energy'] - ref['energy']) <= 5e-4 * abs(ref['energy'])
    # pure elements (one species in the list)
    fe = _rattled('Fe', 3.6, (3, 3, 3), 4)
    _calc_compare(EamFsNN(['Fe'], custom_potentials='msah11'), fe, pot, 'fs', 5.3)
    # analytic Hessian of the msah11 kinds (second-order dual numbers) vs oracle autograd
    with precision_scope('high'):
        nn = EamFsNN(['Al', 'Fe'], custom_potentials='msah11',
                     export_properties=('energy', 'forces', 'hessian'))
        from tensoralloy_b200.transformer import UniversalTransformer
        nn.attach_transformer(UniversalTransformer(['Al', 'Fe'], rcut=6.5))
        base = bulk_fcc('Al', 4.0, (2, 2, 2))
        sym = ['Fe' if x < 0.4 else 'Al' for x in rng.random(len(base))]
        small = Atoms(sym, base.positions + rng.normal(scale=0.08, size=base.positions.shape),
                      base.cell, True)
        calc = TensorAlloyCalculator(nn)
        calc.calculate(small, properties=['energy', 'forces', 'hessian'])
        H = calc.get_hessian(small)
    ref = oeam.eam_evaluate(pot, 'fs', ['Al', 'Fe'], sym, small.positions, small.cell,
                            [1, 1, 1], 6.5, hessian=True)
    n = len(small)
    assert np.abs(H - ref['hessian'].reshape(3 * n, 3 * n)).max() < 1e-8
    assert np.abs(H).max() > 1.0
