"""
Per-structure scalars kept in `atoms.info` -- mirror of tensoralloy/atoms_utils.py:14-69:
the electron temperature (eV) and electron entropy of finite-temperature data and the kinetic
energy; a value may also sit in `info['data']` or `info['key_value_pairs']` (structures read
back from the reference's SQLite store).
"""


def _get(atoms, prop, default):
    info = atoms.info
    if prop in info:
        return info.get(prop)
    if 'data' in info and prop in info['data']:
        return info['data'][prop]
    if 'key_value_pairs' in info and prop in info['key_value_pairs']:
        return info['key_value_pairs'][prop]
    return default


def get_electron_temperature(atoms) -> float:
    return _get(atoms, 'etemperature', 0.0)


def set_electron_temperature(atoms, t: float):
    atoms.info['etemperature'] = t


def get_electron_entropy(atoms) -> float:
    return _get(atoms, 'eentropy', 0.0)


def set_electron_entropy(atoms, eentropy: float):
    atoms.info['eentropy'] = eentropy


def get_kinetic_energy(atoms) -> float:
    return _get(atoms, 'kinetic_energy', 0.0)


def set_kinetic_energy(atoms, ke: float):
    atoms.info['kinetic_energy'] = ke
