"""
Potential-function registry: maps the reference's potential names
(nn/eam/potentials/__init__.py:21-31) to builders of `tab_fn` table entries for
libtab200.  The arithmetic itself lives in csrc/potentials.cuh.
"""
from tensoralloy_b200.nn.eam.potentials.empirical import (
    AgrawalBe, AgSutton90, AlFeMsah11, MishinH, RWGrimes)
from tensoralloy_b200.nn.eam.potentials.zjw04 import (
    Zjw04, Zjw04xc, Zjw04uxc, Zjw04xcp)

available_potentials = {
    'zjw04': Zjw04,
    'zjw04xc': Zjw04xc,
    'zjw04uxc': Zjw04uxc,
    'zjw04xcp': Zjw04xcp,
    'sutton90': AgSutton90,
    'grimes': RWGrimes,
    'mishinh': MishinH,
    'Be/1': AgrawalBe,
    'msah11': AlFeMsah11,
}

def get_potential(name):
    try:
        return available_potentials[name]()
    except KeyError:
        raise ValueError(f"Unknown EAM potential: {name}")
