"""
Unit conversion of dataset labels -- mirror of tensoralloy/io/units.py:17-56.  The
reference evaluates the unit string with `ase.units`; ASE is not installed here, so the
constants it would pick (ase.units, CODATA 2014 -- the default of ASE >= 3.21) are stated.
"""
import re

_UNITS = {
    'eV': 1.0,
    'Hartree': 27.211386024367243,
    'kcal': 2.611447418269555e+22,
    'mol': 6.022140857e+23,
    'Bohr': 0.5291772105638411,
    'Angstrom': 1.0,
    'GPa': 0.006241509125883258,            # = 1 / 160.21766208
    'kbar': 0.1 * 0.006241509125883258,
}
_pattern = re.compile("|".join(re.escape(k) for k in _UNITS))
_allowed = re.compile(r"^[0-9eE+\-*/(). ]*$")


def _parse_comb(comb):
    if not comb:
        return 1.0
    expr = _pattern.sub(lambda m: repr(_UNITS[m.group(0)]), comb)
    if not _allowed.match(expr):
        raise ValueError(f"unknown unit expression: '{comb}'")
    return float(eval(expr, {"__builtins__": {}}, {}))      # arithmetic on numbers only


def get_conversion_units(units):
    """(to_eV, to_eV_Angstrom, to_eV_Ang3) for the keys 'energy', 'forces', 'stress'
    (units.py:31-56); a missing key means the label already has the target unit."""
    units = units or {}
    return (_parse_comb(units.get('energy')), _parse_comb(units.get('forces')),
            _parse_comb(units.get('stress')))
