"""
Ordering rules and small helpers shared by the host-side mirror.
Mirror of the reference's tensoralloy/utils.py (pairing functions :88-161,
k-body term enumeration :210-290, ModeKeys :322-341, Defaults :393-420).
"""
import enum
from itertools import chain
from typing import Dict, List, Tuple

import numpy as np


class ModeKeys(enum.Enum):
    """utils.py:322-341.  LAMMPS / KMC / PRECOMPUTE are referenced by the
    reference but never defined there; they are accepted here as
    aliases of PREDICT-like behaviour so that callers do not crash."""
    TRAIN = 'train'
    EVAL = 'eval'
    PREDICT = 'infer'
    NATIVE = 'native'
    LAMMPS = 'lammps'
    KMC = 'kmc'
    PRECOMPUTE = 'precompute'


class Defaults:
    """utils.py:393-420."""
    rc = 6.0
    k_max = 2
    eta = np.array([0.05, 4.0, 20.0, 80.0])
    omega = np.array([0.0])
    beta = np.array([0.005, ])
    gamma = np.array([1.0, -1.0])
    zeta = np.array([1.0, 4.0])
    n_etas = 4
    n_omegas = 1
    n_betas = 1
    n_gammas = 2
    n_zetas = 2
    cutoff_function = "cosine"
    seed = 611
    variable_moving_average_decay = 0.999
    activation = 'softplus'
    hidden_sizes = [64, 32]
    learning_rate = 0.01


def cantor_pairing(x, y):
    x = np.asarray(x)
    y = np.asarray(y)
    return (x + y) * (x + y + 1) // 2 + y


def _szudzik2(x, y):
    """Szudzik pairing extended to negative integers (utils.py:88-128)."""
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    xx = np.where(x >= 0, 2 * x, -2 * x - 1)
    yy = np.where(y >= 0, 2 * y, -2 * y - 1)
    return np.where(xx >= yy, xx * xx + xx + yy, yy * yy + xx)


def szudzik_pairing(x, *args):
    """utils.py:133-161: fold the pairing over any number of integers / columns."""
    if np.isscalar(x):
        z = int(x)
        for y in args:
            z = int(_szudzik2(z, int(y)))
        return z
    x = np.asarray(x)
    if x.ndim == 1:
        z = x
        for y in args:
            z = _szudzik2(z, np.asarray(y))
        return z
    if x.ndim == 2:
        z = x[:, 0]
        for c in range(1, x.shape[1]):
            z = _szudzik2(z, x[:, c])
        return z
    raise ValueError("Dimension error")


def get_elements_from_kbody_term(kbody_term: str) -> List[str]:
    """'AlCu' -> ['Al', 'Cu']  (utils.py:210-234)."""
    out: List[str] = []
    for ch in kbody_term:
        if ch.isupper():
            out.append(ch)
        else:
            out[-1] += ch
    return out


def get_kbody_terms(elements: List[str], angular=False, symmetric=True
                    ) -> Tuple[List[str], Dict[str, List[str]], List[str]]:
    """Ordered k-body terms (utils.py:237-290).  Defines the feature order of
    every descriptor: per centre element c -> [cc, c-other...] then, if angular,
    c + each sorted pair (j <= k) (or all ordered pairs if not symmetric)."""
    elements = sorted(set(elements))
    n = len(elements)
    per = {e: [f"{e}{e}"] for e in elements}
    for i in range(n):
        for j in range(n):
            if i != j:
                per[elements[i]].append(f"{elements[i]}{elements[j]}")
    if angular:
        for i in range(n):
            c = elements[i]
            for j in range(n):
                ks = range(j, n) if symmetric else range(n)
                for k in ks:
                    if symmetric:
                        suffix = "".join(sorted([elements[j], elements[k]]))
                    else:
                        suffix = f"{elements[j]}{elements[k]}"
                    per[c].append(f"{c}{suffix}")
    all_terms = list(chain(*[per[e] for e in elements]))
    return all_terms, per, elements
