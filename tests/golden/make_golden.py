#!/usr/bin/env python
"""
Copies the reference's golden vectors for the hot path into small fixtures under
tests/golden/ (the reference tree does not exist on the GPU box).  Run in the
build container:  python tests/golden/make_golden.py

Sources (all under /root/reference/test_files):
  lammps/zjw04_Ni.alloy.eam   zjw04 Ni rho / r*phi / F tables written by the
                              reference's EamAlloyNN.export_to_setfl
  lammps/Zhou_AlCu.alloy.eam  golden of nn/eam/tests/test_eam_alloy_nn.py:138-165
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


def main():
    for name in ('zjw04_Ni.alloy.eam', 'Zhou_AlCu.alloy.eam'):
        d = read_setfl(REF / 'lammps' / name)
        np.savez_compressed(OUT / (name.replace('.alloy.eam', '') + '_setfl.npz'), **d)
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
    print('fixtures written to', OUT)


if __name__ == '__main__':
    main()
