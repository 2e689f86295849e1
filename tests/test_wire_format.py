"""The reference's feed-dict wire format (transformer/universal.py:46-233, 851-893): the
vectorised builder of the product (transformer/wire_format.py, driven here with the oracle's
neighbour list -- no GPU) against the loop restatement of the reference (oracle/wire_format.py),
plus the hand-checkable invariants of the reference's layout."""
import numpy as np
import pytest

from oracle import neighbor as onl
from oracle import wire_format as owf
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer
from tensoralloy_b200.transformer import wire_format as wf


def _pd3o2():
    # the reference's AMP fixture geometry (tests/test_utils.py:44-52): Pd3O2, pbc T T F
    pos = np.array([[0., 1., 0.], [1., 2., 1.], [-1., 1., 2.], [1., 3., 2.], [3., 1., 4.]])
    return Atoms('Pd3O2', pos, cell=[4, 6, 8], pbc=[True, True, False])


def _cases():
    ni = bulk_fcc('Ni', 3.52, (2, 2, 2))
    ni.positions += np.random.default_rng(5).normal(scale=0.05, size=ni.positions.shape)
    alloy = bulk_fcc('Ni', 3.6, (2, 2, 2))
    rng = np.random.default_rng(7)
    sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(alloy))]
    alloy = Atoms(sym, alloy.positions + rng.normal(scale=0.08, size=alloy.positions.shape),
                  alloy.cell, True)
    return {'ni': (ni, ['Ni']), 'moni': (alloy, ['Mo', 'Ni']), 'pd3o2': (_pd3o2(), ['O', 'Pd'])}


def _both(atoms, elements, rcut, acut, angular, symmetric):
    clf = UniversalTransformer(elements, rcut=rcut, acut=acut, angular=angular,
                               symmetric=symmetric)
    vap = clf.get_vap_transformer(atoms)
    rmax = max(rcut, acut or 0.0) if angular else rcut
    nl = onl.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, rmax)
    symbols = atoms.get_chemical_symbols()
    n = len(symbols)
    l2g = vap.local_to_gsl_array
    g2l = {int(l2g[k + 1]): k for k in range(n)}
    ref = owf.feed_metadata(symbols, atoms.positions, atoms.cell, nl, clf.elements,
                            clf.kbody_terms_for_element, l2g, g2l, rcut, acut, angular,
                            symmetric)
    with precision_scope('high'):
        feed = clf._feed_dict_from_list(atoms, vap, clf.get_types(atoms), np.asarray(atoms.cell),
                                        atoms.get_volume(), nl[0], nl[1], nl[2], np.float64)
    return clf, vap, ref, feed


@pytest.mark.parametrize('name', ['ni', 'moni', 'pd3o2'])
@pytest.mark.parametrize('angular,symmetric,acut', [(False, True, None), (True, True, None),
                                                   (True, True, 3.5), (True, False, None)])
def test_feed_dict_matches_loop_restatement(name, angular, symmetric, acut):
    atoms, elements = _cases()[name]
    rcut = 4.0 if name != 'pd3o2' else 4.6
    clf, vap, ref, feed = _both(atoms, elements, rcut, acut, angular, symmetric)
    keys = ["g2.v2g_map", "g2.ilist", "g2.jlist", "g2.n1"]
    if angular:
        keys += ["g4.v2g_map", "g4.ilist", "g4.jlist", "g4.klist", "g4.n1", "g4.n2", "g4.n3"]
    for key in keys:
        assert np.array_equal(np.asarray(feed[key]).astype(np.int64),
                              np.asarray(ref[key]).astype(np.int64)), key
    # dtypes and scalars of universal.py:851-893
    assert feed["g2.v2g_map"].dtype == np.int32 and feed["g2.ilist"].dtype == np.int32
    assert feed["g2.n1"].dtype == np.float64
    assert feed["nnl_max"] == feed["g2.v2g_map"][:, 2].max() + 1
    assert feed["n_atoms_vap"] == vap.max_vap_natoms
    assert list(feed["row_splits"]) == [1] + [vap.max_occurs[e] for e in clf.elements]
    assert feed["positions"].shape == (vap.max_vap_natoms, 3)
    if angular:
        assert feed["ij2k_max"] == feed["g4.v2g_map"][:, 3].max() + 1
        assert np.array_equal(feed["g4.n3"], feed["g4.n2"] - feed["g4.n1"])


def test_layout_invariants_of_the_reference():
    """What the reference's scatter_nd relies on: (term, centre, slot) is unique per pair and
    dense from 0; (term, centre, ij slot, ij2k slot) is unique per triple; the number of triples
    is sum n (n - 1) / 2 over the centres."""
    atoms, elements = _cases()['moni']
    clf, vap, ref, feed = _both(atoms, elements, 4.0, None, True, True)
    g2 = feed["g2.v2g_map"]
    assert len({tuple(r[:3]) for r in g2}) == len(g2)
    for key in {tuple(r[:2]) for r in g2}:
        slots = np.sort(g2[(g2[:, 0] == key[0]) & (g2[:, 1] == key[1]), 2])
        assert np.array_equal(slots, np.arange(len(slots)))
    g4 = feed["g4.v2g_map"]
    assert len({tuple(r[:4]) for r in g4}) == len(g4)
    counts = np.bincount(feed["g2.ilist"])
    assert len(g4) == int((counts * (counts - 1) // 2).sum())
    # virtual atom never appears as a centre or neighbour
    assert feed["g2.ilist"].min() >= 1 and feed["g4.klist"].min() >= 1


def test_running_count_helper():
    keys = np.array([5, 3, 5, 5, 3, 9])
    assert list(wf._running_count(keys)) == [0, 0, 1, 2, 1, 0]
