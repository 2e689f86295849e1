"""
Per-structure scalars kept in `atoms.info` -- mirror of tensoralloy/atoms_utils.py:14-69:
the electron temperature (eV) and electron entropy of finite-temperature data and the kinetic
energy (eV).  A value may also sit one level down, in `info['data']` or
`info['key_value_pairs']` (structures read back from the reference's SQLite store); writes
always go to the top level.
"""
_NESTED = ('data', 'key_value_pairs')


def _lookup(atoms, key, default=0.0):
    info = atoms.info
    if key in info:
        return info[key]
    for holder in _NESTED:
        inner = info.get(holder)
        if isinstance(inner, dict) and key in inner:
            return inner[key]
    return default


def _accessors(key):
    def getter(atoms) -> float:
        return _lookup(atoms, key)

    def setter(atoms, value: float):
        atoms.info[key] = value
    getter.__doc__ = f"`{key}` of the structure (0.0 when absent)."
    setter.__doc__ = f"Store `{key}` in `atoms.info`."
    return getter, setter


get_electron_temperature, set_electron_temperature = _accessors('etemperature')
get_electron_entropy, set_electron_entropy = _accessors('eentropy')
get_kinetic_energy, set_kinetic_energy = _accessors('kinetic_energy')
