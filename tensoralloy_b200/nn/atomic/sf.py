"""
SymmetryFunction descriptor -- mirror of the reference's
tensoralloy/nn/atomic/sf.py:25-215 (constructor signature :31-38, `as_dict`
:59-69).  The G2 / G4 arithmetic runs on the GPU (csrc/sf.cu); this class carries
the hyper-parameters and defines their ORDER, which fixes the feature layout:
sklearn's ParameterGrid iterates sorted keys with the last key fastest
(sf.py:47-51) -> radial: eta outer, omega inner; angular: beta outer, gamma,
zeta inner.
"""
import numpy as np


class SymmetryFunction:
    def __init__(self, elements, eta=np.array([0.05, 4.0, 20.0, 80.0]),
                 omega=np.asarray([0.0]), beta=np.asarray([0.005]),
                 gamma=np.asarray([1.0, -1.0]), zeta=np.asarray([1.0, 4.0]),
                 cutoff_function="cosine"):
        self._elements = sorted(list(elements))
        self._eta = np.asarray(eta, dtype=float)
        self._omega = np.asarray(omega, dtype=float)
        self._gamma = np.asarray(gamma, dtype=float)
        self._zeta = np.asarray(zeta, dtype=float)
        self._beta = np.asarray(beta, dtype=float)
        self._cutoff_function = cutoff_function
        self._radial_parameters = [
            {'eta': e, 'omega': o} for e in self._eta for o in self._omega]
        self._angular_parameters = [
            {'beta': b, 'gamma': g, 'zeta': z}
            for b in self._beta for g in self._gamma for z in self._zeta]

    name = property(lambda self: "SF")
    elements = property(lambda self: self._elements)
    cutoff_function = property(lambda self: self._cutoff_function)
    radial_parameters = property(lambda self: self._radial_parameters)
    angular_parameters = property(lambda self: self._angular_parameters)

    def as_dict(self):
        return {"class": self.__class__.__name__, "elements": self._elements,
                "eta": self._eta.tolist(), "omega": self._omega.tolist(),
                "gamma": self._gamma.tolist(), "zeta": self._zeta.tolist(),
                "beta": self._beta.tolist(),
                "cutoff_function": self._cutoff_function}

    def radial_kind(self):
        return 'sf'

    def moments(self):
        return (0,)

    def radial_sets(self):
        return [(p['eta'], p['omega']) for p in self._radial_parameters]

    def angular_sets(self):
        return [(p['beta'], p['gamma'], p['zeta']) for p in self._angular_parameters]

    def dimension(self, angular):
        n = len(self._elements)
        d = n * len(self._radial_parameters)
        if angular:
            d += n * (n + 1) // 2 * len(self._angular_parameters)
        return d
