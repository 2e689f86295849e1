#!/usr/bin/env python
"""
Copies the reference's golden vectors for the hot path into small fixtures under
tests/golden/ (the reference tree does not exist on the GPU box).  Run in the
build container:  python tests/golden/make_golden.py

Sources (all under /root/reference/test_files):
  lammps/zjw04_Ni.alloy.eam   zjw04 Ni rho / r*phi / F tables written by the
                              reference's EamAlloyNN.export_to_setfl
  lammps/Zhou_AlCu.alloy.eam  golden of nn/eam/tests/test_eam_alloy_nn.py:138-165
  lammps/Mendelev_Al_Fe.fs.eam  golden of nn/eam/tests/test_eam_fs_nn.py:60-83 (msah11)
  lammps/Ag.funcfl.eam        golden of nn/eam/potentials/tests/test_sutton90.py:47-57
  lammps/MoNi_Zhou04.eam.alloy  independent Mo / Ni / Mo-Ni tables (test_eam_alloy_nn.py:207)
  amp_Pd3O2.npz               golden of nn/atomic/tests/test_sf.py:666-691
  crystals/Ni_fc2.npy         golden of nn/constraint/tests/test_fc2.py:30-54
  Be_liquid_4000K_TS.extxyz   config-2 geometry (3 x 128 Be atoms)
  models/Mo.zhou04.pb         frozen GraphDef exported by the reference (model-file layout)
"""
import shutil
from pathlib import Path

import numpy as np

REF = Path('/root/reference/test_files')
OUT = Path(__file__).resolve().parent


def read_setfl(path):
    lines = path.read_text().split('\n')
    elements = lines[3].split()[1:]
    nrho, drho, nr, dr, rc = lines[4].split()
    nrho, nr = int(nrho), int(nr)
    drho, dr, rc = float(drho), float(dr), float(rc)
    vals = []
    heads = []
    k = 5
    body = []
    # element blocks: header line + nrho + nr numbers (possibly several per line)
    tokens = []
    for line in lines[5:]:
        tokens.extend(line.split())
    pos = 0
    out = {'elements': np.array(elements), 'nrho': nrho, 'drho': drho, 'nr': nr,
           'dr': dr, 'rc': rc}
    for el in elements:
        pos += 4                       # Z mass a0 lattice
        F = np.array(tokens[pos:pos + nrho], dtype=np.float64)
        pos += nrho
        rho = np.array(tokens[pos:pos + nr], dtype=np.float64)
        pos += nr
        out[f'F_{el}'] = F
        out[f'rho_{el}'] = rho
    for i, a in enumerate(elements):
        for b in elements[:i + 1]:
            out[f'rphi_{a}{b}'] = np.array(tokens[pos:pos + nr], dtype=np.float64)
            pos += nr
    return out


def read_extxyz(path):
    lines = path.read_text().split('\n')
    frames = []
    k = 0
    while k < len(lines) and lines[k].strip():
        n = int(lines[k])
        header = lines[k + 1]
        lat = header.split('Lattice="')[1].split('"')[0]
        cell = np.array(lat.split(), dtype=np.float64).reshape(3, 3)
        rows = [l.split() for l in lines[k + 2:k + 2 + n]]
        sym = [r[0] for r in rows]
        pos = np.array([r[1:4] for r in rows], dtype=np.float64)
        frames.append((sym, pos, cell, header))
        k += 2 + n
    return frames


def copy_text_fixtures():
    """Labelled-structure files of the reference's reader tests (io/tests/test_read.py:19-96,
    test_units.py:41-48): whole files where small, leading frames otherwise."""
    def head_frames(name, n_frames, out):
        lines = (REF / name).read_text().split('\n')
        k = 0
        for _ in range(n_frames):
            k += int(lines[k].split()[0]) + 2
        (OUT / out).write_text('\n'.join(lines[:k]) + '\n')
    head_frames('B28.xyz', 2, 'B28_2frames.xyz')
    head_frames('Be_liquid_4000K_TS.extxyz', 1, 'Be_liquid_4000K_1frame.extxyz')
    shutil.copy(REF / 'examples.extxyz', OUT / 'examples.extxyz')
    shutil.copy(REF / 'snap_Ni_id11.extxyz', OUT / 'snap_Ni_id11.extxyz')
    shutil.copy(REF / 'Pu8.stepmax.xyz', OUT / 'Pu8.stepmax.xyz')
    # vasprun.xml files of io/tests/test_vasp.py:21-39 without the blocks no reader here uses
    # (eigenvalues, dos, k-point lists, primitive cell), gzip-compressed
    import gzip
    import xml.etree.ElementTree as ET
    for name in ('Be_hcp_4000K_vasprun.xml', 'Be_md_vasprun.xml'):
        root = ET.parse(REF / name).getroot()
        for parent in [root] + root.findall('calculation'):
            for tag in ('eigenvalues', 'dos', 'kpoints', 'primitive_cell', 'projected'):
                for child in parent.findall(tag):
                    parent.remove(child)
        with gzip.GzipFile(OUT / (name + '.gz'), 'wb', mtime=0) as fp:
            fp.write(ET.tostring(root))


def copy_databases():
    """ASE SQLite databases of the reference (tests/test_neighbor.py:20-36 reads qm7m.db;
    snap-Ni.db's metadata record holds the neighbour maxima the reference computed for it).
    qm7m.db whole (3 structures); of snap-Ni.db the structures that attain the recorded maxima
    (ids 99, 264, 266, 290: found with the oracle's list over all 461 structures, every recorded
    maximum reproduced) plus three more, re-written with this package's writer, metadata kept."""
    import os
    import sys
    sys.path.insert(0, str(OUT.parents[1]))
    from tensoralloy_b200.io.sqlite import CoreDatabase
    shutil.copy(REF / 'datasets' / 'qm7m' / 'qm7m.db', OUT / 'qm7m.db')
    os.chmod(OUT / 'qm7m.db', 0o644)
    tmp = OUT / '_snap_full.db'
    shutil.copy(REF.parent / 'tensoralloy' / 'data' / 'datasets' / 'snap-Ni.db', tmp)
    os.chmod(tmp, 0o644)
    full = CoreDatabase(tmp)
    sub_path = OUT / 'snap_Ni_subset.db'
    if sub_path.exists():
        sub_path.unlink()
    sub = CoreDatabase(sub_path)
    ids = [1, 11, 99, 150, 264, 266, 290]
    for k in ids:
        a = full.get_atoms(id=k, add_additional_information=True)
        sub.write(a, key_value_pairs=a.info.get('key_value_pairs'), data=a.info.get('data'))
    md = full.metadata
    md['subset_of'] = {'file': 'tensoralloy/data/datasets/snap-Ni.db', 'ids': ids}
    sub.metadata = md
    sub.close()
    full.close()
    tmp.unlink()


def main():
    copy_databases()
    for name in ('zjw04_Ni.alloy.eam', 'Zhou_AlCu.alloy.eam'):
        d = read_setfl(REF / 'lammps' / name)
        np.savez_compressed(OUT / (name.replace('.alloy.eam', '') + '_setfl.npz'), **d)
    # INDEPENDENT tabulation (NIST interatomic potentials repository, generated by Zhou's own
    # Zhou04_create_v2.f): golden of nn/eam/tests/test_eam_alloy_nn.py:207-217 (via LAMMPS)
    d = read_setfl(REF / 'lammps' / 'MoNi_Zhou04.eam.alloy')
    np.savez_compressed(OUT / 'MoNi_Zhou04_setfl.npz', **d)
    # funcfl table of the Sutton-Chen Ag potential: golden of
    # nn/eam/potentials/tests/test_sutton90.py:47-57 (every 10th knot keeps the fixture small)
    lines = (REF / 'lammps' / 'Ag.funcfl.eam').read_text().split('\n')
    nrho, drho, nr, dr, rc = lines[2].split()
    nrho, nr = int(nrho), int(nr)
    tok = ' '.join(lines[3:]).split()
    np.savez_compressed(
        OUT / 'Ag_funcfl.npz', drho=float(drho) * 10, dr=float(dr) * 10, rc=float(rc),
        F=np.array(tok[:nrho], dtype=np.float64)[::10],
        Z=np.array(tok[nrho:nrho + nr], dtype=np.float64)[::10],
        rho=np.array(tok[nrho + nr:nrho + 2 * nr], dtype=np.float64)[::10])
    # eam/fs table of the Mendelev Al-Fe potential: golden of
    # nn/eam/tests/test_eam_fs_nn.py:60-83 (delta 1e-8); every 10th knot
    lines = (REF / 'lammps' / 'Mendelev_Al_Fe.fs.eam').read_text().split('\n')
    nrho, drho, nr, dr, rc = lines[4].split()
    nrho, nr = int(nrho), int(nr)
    tok = ' '.join(lines[5:]).split()
    pos, fs = 0, {'drho': float(drho) * 10, 'dr': float(dr) * 10, 'rc': float(rc)}
    for el in ('Al', 'Fe'):
        pos += 4
        fs['F_' + el] = np.array(tok[pos:pos + nrho], dtype=np.float64)[::10]
        pos += nrho
        for b in ('Al', 'Fe'):       # density at a centre `el` from a neighbour `b`
            fs[f'rho_{el}{b}'] = np.array(tok[pos:pos + nr], dtype=np.float64)[::10]
            pos += nr
    for key in ('AlAl', 'FeAl', 'FeFe'):
        fs['rphi_' + key] = np.array(tok[pos:pos + nr], dtype=np.float64)[::10]
        pos += nr
    np.savez_compressed(OUT / 'Mendelev_AlFe_fs.npz', **fs)
    shutil.copy(REF / 'amp_Pd3O2.npz', OUT / 'amp_Pd3O2.npz')
    shutil.copy(REF / 'crystals' / 'Ni_fc2.npy', OUT / 'Ni_fc2.npy')
    # a frozen model exported by the reference (legacy metadata layout)
    shutil.copy(REF / 'models' / 'Mo.zhou04.pb', OUT / 'Mo.zhou04.pb')
    frames = read_extxyz(REF / 'Be_liquid_4000K_TS.extxyz')
    np.savez_compressed(
        OUT / 'Be_liquid_4000K.npz',
        positions=np.array([f[1] for f in frames]),
        cells=np.array([f[2] for f in frames]),
        symbols=np.array(frames[0][0]),
        headers=np.array([f[3] for f in frames]))
    copy_text_fixtures()
    print('fixtures written to', OUT)


if __name__ == '__main__':
    main()
