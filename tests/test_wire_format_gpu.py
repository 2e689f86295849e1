"""`UniversalTransformer.get_np_feed_dict` end to end (GPU neighbour list -> the reference's
feed-dict wire format, universal.py:851-893) against the loop restatement of the reference on the
oracle's neighbour list: every g2.* and g4.* array bit-exact."""
import numpy as np
import pytest

from oracle import neighbor as onl
from oracle import wire_format as owf
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer
from test_wire_format import _cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['ni', 'moni', 'pd3o2'])
@pytest.mark.parametrize('angular,acut', [(False, None), (True, None), (True, 3.5)])
def test_get_np_feed_dict_bit_exact(name, angular, acut):
    atoms, elements = _cases()[name]
    rcut = 4.0 if name != 'pd3o2' else 4.6
    clf = UniversalTransformer(elements, rcut=rcut, acut=acut, angular=angular)
    with precision_scope('high'):
        feed = clf.get_np_feed_dict(atoms)
    vap = clf.get_vap_transformer(atoms)
    rmax = max(rcut, acut or 0.0) if angular else rcut
    nl = onl.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, rmax)
    symbols = atoms.get_chemical_symbols()
    l2g = vap.local_to_gsl_array
    g2l = {int(l2g[k + 1]): k for k in range(len(symbols))}
    ref = owf.feed_metadata(symbols, atoms.positions, atoms.cell, nl, clf.elements,
                            clf.kbody_terms_for_element, l2g, g2l, rcut, acut, angular, True)
    for key, val in ref.items():
        assert np.array_equal(np.asarray(feed[key]).astype(np.int64),
                              np.asarray(val).astype(np.int64)), key
    assert feed["nnl_max"] == ref["g2.v2g_map"][:, 2].max() + 1
    if angular:
        assert feed["ij2k_max"] == ref["g4.v2g_map"][:, 3].max() + 1
    # positions in GSL order with the virtual atom in row 0 (vap.py:26-56)
    assert np.array_equal(feed["positions"][0], np.zeros(3))
    assert feed["positions"].shape[0] == feed["n_atoms_vap"]


@pytest.mark.parametrize('angular', [False, True])
def test_tfrecord_example_from_gpu_list(angular):
    """`BatchUniversalTransformer.encode` with the GPU list builder writes the same record,
    byte for byte, as with the oracle's neighbour list (universal.py:1219-1230)."""
    from tensoralloy_b200.transformer import BatchUniversalTransformer
    atoms, _ = _cases()['moni']
    atoms.info.update(energy=-3.25, forces=np.full((len(atoms), 3), 0.125))
    rc = 4.0
    nl = onl.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, rc)[:3]
    clf = BatchUniversalTransformer({'Mo': 20, 'Ni': 30}, rcut=rc, angular=angular,
                                    nij_max=len(nl[0]) + 7,
                                    nijk_max=60000 if angular else None)
    with precision_scope('high'):
        a = clf.encode(atoms).SerializeToString()
        b = clf.encode(atoms, neighbor_list=nl).SerializeToString()
        assert a == b
        dec = clf.decode_protobuf(a)
    assert dec['energy'] == -3.25 and int(dec['g2.v2g_map'][:, 5].sum()) == len(nl[0])
