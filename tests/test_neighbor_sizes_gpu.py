"""find_neighbor_size_of_atoms (reference neighbor.py:50-146) and the 'elastic'
property (nn/constraint/elastic.py:24-91) against the reference's own asserted
values."""
import numpy as np
import pytest

from tensoralloy_b200.atoms import Atoms
from tensoralloy_b200.neighbor import NeighborProperty, find_neighbor_size_of_atoms

pytestmark = pytest.mark.gpu


def test_find_sizes_reference_values():
    # reference tests/test_neighbor.py:20-36 uses qm7m.db ids 2 and 3 (CH4, C2H6;
    # all atoms within rc = 6.5): nij 20 / 56, nnl 4 / 6, nijk 0 / 168
    ch4 = Atoms('CH4', positions=[[0, 0, 0], [0.63, 0.63, 0.63], [-0.63, -0.63, 0.63],
                                  [-0.63, 0.63, -0.63], [0.63, -0.63, -0.63]], pbc=False)
    size = find_neighbor_size_of_atoms(ch4, 6.5, find_nijk=False)
    assert (size.nij, size.nijk, size.nnl) == (20, 0, 4)
    c2h6 = Atoms('C2H6', positions=[[0, 0, 0.77], [0, 0, -0.77],
                                    [1.02, 0, 1.16], [-0.51, 0.88, 1.16], [-0.51, -0.88, 1.16],
                                    [-1.02, 0, -1.16], [0.51, 0.88, -1.16], [0.51, -0.88, -1.16]],
                 pbc=False)
    size = find_neighbor_size_of_atoms(c2h6, 6.5, find_nijk=True)
    assert (size.nij, size.nijk, size.nnl) == (56, 168, 6)
    assert size[NeighborProperty.nij] == 56 and size['nnl'] == 6
    # ij2k: C centre, neighbour H -> 5 other H
    size = find_neighbor_size_of_atoms(c2h6, 6.5, find_ij2k=True)
    assert size.ij2k == 6        # H centre: neighbour C, others of species H = 5; C: 6 H
    # translated far from the origin / negative coordinates: same sizes
    far = Atoms('C2H6', positions=c2h6.positions - 37.3, pbc=False)
    assert find_neighbor_size_of_atoms(far, 6.5, find_nijk=True).nij == 56


def test_elastic_property_single_atom_cell():
    """'elastic' with the reference's definition on the fcc Ni primitive cell:
    C11 246.61, C12 147.15, C44 124.72 GPa (tests/test_calculator.py:101-108)."""
    from tensoralloy_b200.calculator import TensorAlloyCalculator
    from tensoralloy_b200.nn.eam import EamAlloyNN
    from tensoralloy_b200.precision import precision_scope
    from tensoralloy_b200.transformer import UniversalTransformer
    a = 3.52
    cell = 0.5 * a * np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]], dtype=float)
    atoms = Atoms(['Ni'], [[0, 0, 0]], cell, True)
    with precision_scope('high'):
        nn = EamAlloyNN(['Ni'], custom_potentials='zjw04',
                        export_properties=['energy', 'forces', 'stress', 'elastic'])
        nn.attach_transformer(UniversalTransformer(['Ni'], rcut=6.0))
        calc = TensorAlloyCalculator(nn)
        C = calc.get_elastic_constant_tensor(atoms)
    assert abs(C[0, 0] - 246.61) < 0.02
    assert abs(C[0, 1] - 147.15) < 0.02
    assert abs(C[3, 3] - 124.72) < 0.02
    assert np.abs(C - C.T).max() < 1e-8


def test_analytic_elastic_op_equals_the_differenced_virial():
    """`tab_eam_elastic` (closed-form d virial / d h, nn/constraint/elastic.py:24-91) against
    the central difference of the GPU virial over the nine lattice components (the definition
    itself, BasicNN._elastic): primitive cell, a rattled 4-atom cell (only image pairs move
    with h) and a two-species FS-like alloy cell."""
    from tensoralloy_b200.nn.basic import BasicNN
    from tensoralloy_b200.nn.eam import EamAlloyNN
    from tensoralloy_b200.precision import precision_scope
    from tensoralloy_b200.transformer import UniversalTransformer
    a = 3.52
    prim = Atoms(['Ni'], [[0, 0, 0]],
                 0.5 * a * np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]], dtype=float), True)
    rng = np.random.default_rng(3)
    conv = Atoms(['Ni'] * 4, np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]]) * a +
                 rng.normal(scale=0.03, size=(4, 3)), np.eye(3) * a, True)
    alloy = Atoms(['Mo', 'Ni', 'Ni', 'Mo'], conv.positions * (3.6 / a), np.eye(3) * 3.6, True)
    for atoms, elements in ((prim, ['Ni']), (conv, ['Ni']), (alloy, ['Mo', 'Ni'])):
        with precision_scope('high'):
            nn = EamAlloyNN(elements, custom_potentials='zjw04',
                            export_properties=['energy', 'forces', 'stress', 'elastic'])
            clf = UniversalTransformer(elements, rcut=6.0)
            nn.attach_transformer(clf)
            feats = clf.get_constant_features(atoms)
            C = nn._elastic(feats)
            C_fd = BasicNN._elastic(nn, feats)
        assert np.abs(C - C_fd).max() < 2e-5 * max(1.0, np.abs(C_fd).max()), (C, C_fd)
