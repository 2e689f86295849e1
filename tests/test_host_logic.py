"""CPU tests of the host-side mirror: ordering rules (reference tests/test_utils.py,
transformer/tests/test_vap.py), the Atoms stand-in, and that the C-ABI library
loads and exports every symbol include/tab200.h declares (no compute calls)."""
import os
import re
from collections import Counter

import numpy as np
import pytest

from tensoralloy_b200 import utils
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.transformer.vap import VirtualAtomMap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kbody_terms_order():
    # reference tests/test_utils.py: order of terms defines the feature order
    all_terms, per, els = utils.get_kbody_terms(['Ni', 'Al'], angular=False)
    assert els == ['Al', 'Ni']
    assert all_terms == ['AlAl', 'AlNi', 'NiNi', 'NiAl']
    all_terms, per, els = utils.get_kbody_terms(['A', 'B', 'C'], angular=True)
    assert per['A'] == ['AA', 'AB', 'AC', 'AAA', 'AAB', 'AAC', 'ABB', 'ABC', 'ACC']
    assert per['B'][:3] == ['BB', 'BA', 'BC']
    _, per, _ = utils.get_kbody_terms(['A', 'B'], angular=True, symmetric=False)
    assert per['A'] == ['AA', 'AB', 'AAA', 'AAB', 'ABA', 'ABB']
    assert utils.get_elements_from_kbody_term('AlCuNi') == ['Al', 'Cu', 'Ni']


def test_pairing_functions():
    # Szudzik pairing with negative support (utils.py:88-161)
    seen = set()
    for x in range(-3, 4):
        for y in range(-3, 4):
            z = utils.szudzik_pairing(x, y)
            assert z not in seen
            seen.add(z)
    xs = np.array([0, 1, -1, 5, -7])
    ys = np.array([0, -1, 1, 3, 2])
    vec = utils.szudzik_pairing(xs, ys)
    for k in range(len(xs)):
        assert vec[k] == utils.szudzik_pairing(int(xs[k]), int(ys[k]))
    assert utils.szudzik_pairing(3, 4, 5) == utils.szudzik_pairing(
        utils.szudzik_pairing(3, 4), 5)
    assert utils.cantor_pairing(3, 4) == (7 * 8) // 2 + 4


def test_vap_matches_reference_semantics():
    # transformer/tests/test_vap.py:24-64 style: Pd3O2 into a Pd4 O5 template
    symbols = ['Pd', 'Pd', 'O', 'Pd', 'O']
    max_occurs = Counter({'Pd': 4, 'O': 5})
    vap = VirtualAtomMap(max_occurs, symbols)
    assert vap.max_vap_natoms == 10
    # sorted elements: O (5 slots: 1..5), Pd (4 slots: 6..9)
    l2g = vap.local_to_gsl_map
    assert [l2g[i] for i in range(6)] == [0, 6, 7, 1, 8, 2]
    g2l = vap.gsl_to_local_map
    assert g2l[6] == 0 and g2l[1] == 2 and g2l.get(3, -1) == -1
    assert vap.atom_masks.tolist() == [False, True, True, False, False, False,
                                       True, True, True, False]
    pos = np.arange(15, dtype=float).reshape(5, 3)
    mapped = vap.map_positions(pos)
    assert mapped.shape == (10, 3)
    assert np.array_equal(mapped[6], pos[0]) and np.array_equal(mapped[1], pos[2])
    assert np.all(mapped[0] == 0) and np.all(mapped[3] == 0)
    back = vap.map_positions(mapped, reverse=True)
    assert np.array_equal(back, pos)
    assert vap.vap_symbols == ['X'] + ['O'] * 5 + ['Pd'] * 4
    # hessian remap
    H = np.random.default_rng(0).random((10, 3, 10, 3))
    h2 = vap.reverse_map_hessian(H)
    assert h2.shape == (15, 15)
    assert h2[0 * 3 + 1, 2 * 3 + 2] == H[6, 1, 1, 2]
    h4 = vap.reverse_map_hessian(H, phonopy_format=True)
    assert h4[3, 4, 0, 2] == H[8, 0, 2, 2]
    with pytest.raises(ValueError):
        vap.reverse_map_hessian(np.zeros((3, 3)))
    single = VirtualAtomMap(Counter({'Ni': 4}), ['Ni'] * 4)
    assert single.is_identity and not vap.is_identity


def test_atoms_stand_in():
    atoms = bulk_fcc('Ni', 3.52, (2, 2, 2))
    assert len(atoms) == 32 and abs(atoms.get_volume() - 7.04 ** 3) < 1e-9
    assert atoms.get_chemical_formula(mode='reduce') == 'Ni32'
    a = Atoms('Pd3O2', positions=np.zeros((5, 3)), cell=[3, 3, 3], pbc=[1, 1, 0])
    assert a.get_chemical_symbols() == ['Pd', 'Pd', 'Pd', 'O', 'O']
    assert a.get_pbc().tolist() == [True, True, False]


def test_library_exports_every_declared_symbol():
    from tensoralloy_b200 import _build, _lib
    _build.build_library()
    header = open(os.path.join(ROOT, 'include', 'tab200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(tab_[a-z0-9_]+)\s*\(', header))
    assert declared, "no declarations parsed"
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in tab200.h but not exported"
    assert declared == set(_lib.EXPORTS)
    assert L.tab_version() >= 100


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under tensoralloy_b200/ may
    import it (a product path routed through it would void parity)."""
    pkg = os.path.join(ROOT, 'tensoralloy_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_atomic_slab_layout_geometry():
    """2 rc halo of the AtomicNN / ADP decomposition (domain_atomic.py): send masks, the
    inner / outer split and the binning frame."""
    import pytest
    from tensoralloy_b200.domain_atomic import AtomicSlabLayout
    lay = AtomicSlabLayout(40.0, 2, 1, 4.0)           # slab [20, 40), halo 8
    assert (lay.lo, lay.hi, lay.halo) == (20.0, 40.0, 8.0)
    x = np.array([20.5, 27.9, 28.1, 31.9, 32.1, 39.9])
    to_left, to_right = lay.send_masks(x)
    assert to_left.tolist() == [True, True, False, False, False, False]
    assert to_right.tolist() == [False, False, False, False, True, True]
    assert lay.shift_to_right == -40.0 and lay.shift_to_left == 0.0   # last rank wraps right
    cell, origin, pbc = lay.frame(10.0, 12.0)
    assert pbc == [0, 1, 1] and origin[0] == 20.0 - 8.0 - 0.5
    assert cell[0, 0] == 20.0 + 16.0 + 1.0 and cell[1, 1] == 10.0
    with pytest.raises(ValueError):
        AtomicSlabLayout(40.0, 8, 0, 4.0)             # width 5 < halo 8
    AtomicSlabLayout(40.0, 1, 0, 30.0)                # a single rank is never too narrow


def test_transformer_element_lookup_table():
    from tensoralloy_b200.transformer import UniversalTransformer
    clf = UniversalTransformer(['Ni', 'Mo'], rcut=5.0)
    lut = clf._z_lut()
    assert lut[42] == 0 and lut[28] == 1 and (np.delete(lut, [28, 42]) == -1).all()
    atoms = Atoms(['Ni', 'Mo', 'Ni'], np.zeros((3, 3)), np.eye(3) * 5, True)
    assert atoms.numbers.tolist() == [28, 42, 28]
    assert lut[atoms.numbers].tolist() == clf.get_types(atoms).tolist() == [1, 0, 1]
    b = atoms.copy()
    b.positions[0, 0] = 1.0
    assert atoms.positions[0, 0] == 0.0 and b.numbers is not atoms.numbers


def test_phonon_masses_and_units():
    from tensoralloy_b200.analysis.phonon import VaspToCm, VaspToTHz, get_masses
    atoms = Atoms(['Ni', 'Be'], np.zeros((2, 3)), np.eye(3) * 5, True)
    assert np.allclose(get_masses(atoms), [58.6934, 9.0121831])
    # sqrt(eV / (amu A^2)) / (2 pi) in THz
    ev, amu = 1.602176634e-19, 1.66053906660e-27
    assert abs(VaspToTHz - np.sqrt(ev / amu) / 1e-10 / (2 * np.pi) / 1e12) < 1e-4
    assert abs(VaspToCm / VaspToTHz - 33.356410) < 1e-6


def test_c_abi_argument_errors_without_a_gpu():
    """Status codes and tab_last_error() of the entry points added for batches, pair
    operators and decomposition: argument / state validation happens before any CUDA call."""
    import ctypes as C
    from tensoralloy_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.tab_nbr_create(C.byref(h)) == 0
    try:
        off = (C.c_int32 * 2)(0, 4)
        cell = (C.c_double * 9)(*np.eye(3).ravel())
        pbc = (C.c_int32 * 3)(1, 1, 1)
        EINVAL, ESTATE = -1, -5
        # no structures / null positions
        assert L.tab_nbr_build_batch(h, 0, off, None, None, cell, pbc, 5.0, None) == EINVAL
        assert L.tab_nbr_build_batch(h, 1, off, None, None, cell, pbc, 5.0, None) == EINVAL
        assert b'tab_nbr_build_batch' in L.tab_last_error()
        # operators on lists that were never built
        assert L.tab_pairs_export(h, None, None, C.c_void_p(8), None) == ESTATE
        assert b'before tab_nbr_build' in L.tab_last_error()
        assert L.tab_pair_forces(h, C.c_void_p(8), None, None, None) == ESTATE
        assert L.tab_pair_jvp(h, C.c_void_p(8), C.c_void_p(8), C.c_void_p(8), None) == ESTATE
        assert L.tab_nbr_batch_size(h) == 0
        assert L.tab_atomic_eval_dd(None, h, 0, None, None, None, None, None, None) == EINVAL
        assert L.tab_eam_eval_dd(None, h, 0, None, None, None, None, None, None) == EINVAL
    finally:
        L.tab_nbr_free(h)


def test_grap_new_mode_layout_and_refusals():
    """grap.py:613, 662-676: new mode holds every moment 0..max per (term, tau);
    legacy mode only the requested ones (grap.py:419-457)."""
    from tensoralloy_b200.nn.atomic import GenericRadialAtomicPotential as Grap
    par = dict(eta=[0.5, 2.0, 8.0], omega=[0.0, 0.0, 0.0])
    legacy = Grap(['Al', 'Be'], 'sf', par, moment_tensors=[0, 2])
    new = Grap(['Al', 'Be'], 'sf', par, moment_tensors=[2], legacy_mode=False)
    assert legacy.moments() == (0, 2) and legacy.dimension() == 2 * 3 * 2
    assert new.moments() == (0, 1, 2) and new.dimension() == 2 * 3 * 3
    assert new.as_dict()["moment_tensors"] == [2] and new.as_dict()["legacy_mode"] is False
    assert legacy.grap_flags() == 0 and new.grap_flags() == 1
    sym = Grap(['Be'], 'sf', par, moment_tensors=3, symmetric=True, legacy_mode=False)
    assert sym.grap_flags() == 3 and sym.moments() == (0, 1, 2, 3)
    assert Grap(['Be'], 'sf', par, moment_tensors=2, symmetric=True).grap_flags() == 0
    with pytest.raises(ValueError, match="legacy mode"):          # grap.py:296-299
        Grap(['Be'], 'nn', {})
    fnn = Grap(['Be', 'W'], 'nn', dict(num_filters=6, hidden_sizes=[8, 8]),
               moment_tensors=2, legacy_mode=False)
    assert fnn.dimension() == 2 * 6 * 3 and fnn.as_dict()["parameters"]["num_filters"] == 6
    with pytest.raises(ValueError, match="h_abck_modifier"):
        Grap(['Be'], 'nn', dict(h_abck_modifier=3), legacy_mode=False)
    with pytest.raises(ValueError, match="moments 0, 1, 2"):
        Grap(['Be'], 'sf', par, moment_tensors=3)                 # legacy stops at 2
    five = Grap(['Be'], 'sf', par, moment_tensors=5, legacy_mode=False)
    assert five.uses_torch_path() and five.moments() == (0, 1, 2, 3, 4, 5)
    assert not sym.uses_torch_path() and not legacy.uses_torch_path()
    with pytest.raises(ValueError, match="moments 0, 1, 2"):
        Grap(['Be'], 'sf', par, moment_tensors=6, legacy_mode=False)


def test_batch_universal_transformer_mirror():
    """transformer/universal.py:921-1048, 1112-1144: constructor, as_dict, sizes,
    as_descriptor_transformer, row splits; host-side refusals."""
    from collections import Counter
    from tensoralloy_b200.atoms import Atoms
    from tensoralloy_b200.transformer import BatchUniversalTransformer, UniversalTransformer
    blf = BatchUniversalTransformer(Counter({'W': 8, 'Be': 8}), rcut=5.0, acut=4.0,
                                    angular=True, nij_max=500, nijk_max=4000, nnl_max=40,
                                    ij2k_max=10, batch_size=4, use_stress=True)
    assert blf.elements == ['Be', 'W'] and blf.max_n_atoms == 16
    d = blf.as_dict()
    assert d['class'] == 'BatchUniversalTransformer' and d['max_occurs'] == {'W': 8, 'Be': 8}
    assert (d['nij_max'], d['nijk_max'], d['nnl_max'], d['ij2k_max'], d['batch_size']) == \
        (500, 4000, 40, 10, 4)
    assert d['use_forces'] is True and d['use_stress'] is True
    assert blf.get_row_split_sizes(None) == [1, 8, 8] and blf.get_row_split_axis() == 2
    assert blf.get_g_shape(None) == [4, blf.max_nr_terms, 17, 40, 1]
    assert blf.get_g_shape(None, angular=True) == [4, blf.max_na_terms, 17, 40, 10]
    clf = blf.as_descriptor_transformer()
    assert type(clf) is UniversalTransformer
    assert clf.as_dict() == UniversalTransformer(['Be', 'W'], rcut=5.0, acut=4.0,
                                                 angular=True).as_dict()
    assert clf.kbody_terms_for_element == blf.kbody_terms_for_element
    ok = Atoms(['Be'] * 8 + ['W'] * 3, np.random.default_rng(0).random((11, 3)) * 5,
               np.eye(3) * 5, True)
    blf.check_occurs([ok, ok])
    big = Atoms(['Be'] * 9, np.random.default_rng(0).random((9, 3)) * 5, np.eye(3) * 5, True)
    with pytest.raises(ValueError, match="max_occurs"):
        blf.check_occurs([ok, big])
    with pytest.raises(ValueError, match="batch_size"):
        blf.get_batch_features([ok] * 5)
    # the TFRecord codec is covered by tests/test_tfrecord.py
    assert callable(blf.encode) and callable(blf.decode_protobuf)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference`: ONE JSON line with the keys the driver reads; the oracle
    port on a small bounded sample (the default sample is 16^3 cells)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '3', '--ref-cells', '6'],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "atom-evals/s" and d["value"] > 0
    assert d["metric"].startswith("atom-evals/sec (E+F+virial)") and d["higher_is_better"]
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 3
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "1000188 atoms" in d["config"]["workload"] and "sample" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_grap_parameter_grid_and_serialization():
    """nn/atomic/tests/test_grap.py:25-46 (`test_gen_algorithm`, `test_serialization`) and
    grap.py:45-72: 'pair' zips the lists, 'cross' is sklearn's ParameterGrid (keys sorted, last
    key fastest); `as_dict` round-trips through the constructor."""
    from tensoralloy_b200.nn.atomic import GenericRadialAtomicPotential as Grap
    rl, pl = [1.0, 2.0, 3.0, 4.0, 5.0], [2.0, 3.0, 4.0, 5.0, 6.0]
    pair = Grap(['Be'], 'pexp', dict(rl=rl, pl=pl), param_space_method='pair')
    assert len(pair.grid) == 5
    for i, row in enumerate(pair.grid):
        assert row['rl'] == rl[i] and row['pl'] == pl[i]
    assert pair.radial_sets() == list(zip(rl, pl))
    with pytest.raises(ValueError, match="same length"):
        Grap(['Be'], 'pexp', dict(rl=rl, pl=pl[:3]), param_space_method='pair')
    cross = Grap(['Be'], 'sf', dict(eta=[1.0, 2.0], omega=[0.0, 0.5, 1.0]),
                 param_space_method='cross')
    assert [(r['eta'], r['omega']) for r in cross.grid] == \
        [(1.0, 0.0), (1.0, 0.5), (1.0, 1.0), (2.0, 0.0), (2.0, 0.5), (2.0, 1.0)]
    d = cross.as_dict()
    assert d["@class"] == "GenericRadialAtomicPotential" and d["algorithm"] == "sf"
    again = Grap(**{k: v for k, v in d.items() if not k.startswith('@')})
    assert again.as_dict() == d and again.grid == cross.grid
    with pytest.raises(ValueError, match="required"):
        Grap(['Be'], 'morse', dict(D=[1.0], gamma=[1.0]))            # r0 missing
    with pytest.raises(ValueError, match="param_space_method"):
        Grap(['Be'], 'sf', dict(eta=[1.0], omega=[0.0]), param_space_method='zip')


def test_atoms_utils_lookup_order():
    """atoms_utils.py:14-69: info, then info['data'], then info['key_value_pairs']."""
    from tensoralloy_b200 import atoms_utils as au
    from tensoralloy_b200.atoms import Atoms
    atoms = Atoms(['Be'], [[0, 0, 0]], np.eye(3) * 3, True)
    assert au.get_electron_temperature(atoms) == 0.0 and au.get_electron_entropy(atoms) == 0.0
    atoms.info['key_value_pairs'] = {'etemperature': 0.3, 'eentropy': 2.0}
    assert au.get_electron_temperature(atoms) == 0.3 and au.get_electron_entropy(atoms) == 2.0
    atoms.info['data'] = {'etemperature': 0.2}
    assert au.get_electron_temperature(atoms) == 0.2
    au.set_electron_temperature(atoms, 0.1)
    au.set_electron_entropy(atoms, 5.0)
    au.set_kinetic_energy(atoms, 1.5)
    assert (au.get_electron_temperature(atoms), au.get_electron_entropy(atoms),
            au.get_kinetic_energy(atoms)) == (0.1, 5.0, 1.5)


def test_ctypes_mirrors_match_the_header_layouts(tmp_path):
    """Every struct that crosses the C ABI: sizeof and the offset of every field as gcc lays
    out include/tab200.h must equal the ctypes mirror in _lib.py."""
    import ctypes as C
    import shutil
    import subprocess
    from tensoralloy_b200 import _lib
    if shutil.which('gcc') is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = (('tab_fn', _lib.TabFn), ('tab_sf_desc', _lib.TabSfDesc),
             ('tab_mlp_desc', _lib.TabMlpDesc))
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "tab200.h"', 'int main(){']
    for cname, cls in pairs:
        prog.append(f'printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for field, _ in cls._fields_:
            prog.append(f'printf("{cname} {field} %zu\\n", offsetof({cname}, {field}));')
    prog.append('return 0;}')
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(prog))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', os.path.join(root, 'include'), str(src), '-o', str(exe)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = {(a, b): int(c) for a, b, c in (ln.split() for ln in out.splitlines())}
    for cname, cls in pairs:
        assert got[(cname, 'sizeof')] == C.sizeof(cls), cname
        for field, _ in cls._fields_:
            assert got[(cname, field)] == getattr(cls, field).offset, (cname, field)


def test_ctypes_prototypes_have_the_header_arity():
    """Every entry point whose argtypes are declared in _lib.py takes as many arguments as
    its prototype in include/tab200.h (a silent mismatch would corrupt the call frame)."""
    from tensoralloy_b200 import _build, _lib
    _build.build_library()
    header = open(os.path.join(ROOT, 'include', 'tab200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    protos = dict(re.findall(r'\b(tab_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', header))
    L = _lib.lib()
    checked = 0
    for name, args in protos.items():
        fn = getattr(L, name)
        if fn.argtypes is None:
            continue
        args = args.strip()
        arity = 0 if args in ('', 'void') else len(args.split(','))
        assert len(fn.argtypes) == arity, (name, len(fn.argtypes), arity)
        checked += 1
    assert checked >= 30, checked
