"""
GenericRadialAtomicPotential (GRAP) descriptor, legacy mode -- mirror of the
reference's tensoralloy/nn/atomic/grap.py (constructor :236-269, `as_dict`
:306-322, algorithms :121-234, legacy descriptor functions :384-466).

For every radial k-body term `c-x` of a centre of element c, every parameter set
tau of the chosen algorithm f_tau(r) and every requested multipole moment:

    m = 0   sum_j f(r_ij) fc(r_ij)
    m = 1   sum_a  ( sum_j f fc d_a / r )^2              a in x, y, z
    m = 2   sum_ab ( sum_j f fc d_a d_b / r^2 )^2        all 9 (a, b)
    m = 3   sum_abc ( sum_j f fc d_a d_b d_c / r^3 )^2   all 27 (a, b, c); new mode only

laid out per term as [tau][moment] (grap.py:419-457), terms in
`kbody_terms_for_element[c]` order.  The arithmetic runs on the GPU
(csrc/sf.cu: rad_fn, k_sf_forward, k_sf_backward).

New mode (`legacy_mode=False`, the "T_dm / M_dnac" formulation, grap.py:596-680):
the descriptor holds EVERY moment 0..max(moment_tensors) per (term, tau)
(`ndims = (max_moment + 1) * len(algo) * n_elements`, grap.py:613) and, with the
unsymmetric multiplicity tensor (grap.py:471-496: 1 | 1 1 1 | 1 2 2 1 2 1 over the
unique index pairs), the same sums as above -- the reference's own test states the
equality (nn/atomic/tests/test_grap.py:152-200).  It is served by the same kernels
with the moment list widened to 0..max and two flags of the C ABI
(include/tab200.h: TAB_GRAP_SIGNED_SQRT_M0 -- the m = 0 entry is
sign(P) sqrt(P^2 + 1e-16), grap.py:667-676; TAB_GRAP_TRACELESS -- `symmetric=True`,
grap.py:485-494).  Moment 3 (ten unique third-order sums, multiplicities
1 3 3 3 6 3 1 3 3 1) exists in new mode only, as in the reference (its legacy loop
stops at 2, grap.py:434-457).  The trainable `nn` algorithm (a filter network r -> R^K,
grap.py:211-234) is served by nn/atomic/grap_nn.py (torch over the library's pair vectors
and pair-force op), not by the descriptor kernels, and so are moments 4 and 5 (full 3^m
moment tensors with unit weights for every moment, no traceless form: `get_moment_tensor` /
`get_T_dm`, grap.py:537-594).
"""
import numpy as np

ALGORITHMS = {            # name -> required keys, in the order the kernel expects
    'sf': ('eta', 'omega'),
    'morse': ('D', 'gamma', 'r0'),
    'density': ('A', 'beta', 're'),
    'pexp': ('rl', 'pl'),
}


def _parameter_grid(params, keys, method):
    """grap.py:45-72: 'cross' = sklearn ParameterGrid (sorted keys, last key
    fastest), 'pair' = zip."""
    if method == 'pair':
        sizes = {len(params[k]) for k in keys}
        if len(sizes) > 1:
            raise ValueError("Hyperparameters must have the same length for gen:pair")
        return [{k: float(params[k][i]) for k in keys} for i in range(sizes.pop())]
    grid = [{}]
    for k in sorted(keys):
        grid = [dict(g, **{k: float(v)}) for g in grid for v in params[k]]
    return grid


class GenericRadialAtomicPotential:
    def __init__(self, elements, algorithm='sf', parameters=None,
                 param_space_method='pair', moment_tensors=0,
                 cutoff_function='cosine', symmetric=False, legacy_mode=True,
                 h_abck_modifier=None):
        self._algo_nn = None
        if algorithm == 'nn':
            # grap.py:296-301: the filter network exists in new mode only
            if legacy_mode:
                raise ValueError("The NN algorithm cannot be used for GRAP legacy mode")
            from tensoralloy_b200.nn.atomic.grap_nn import NNAlgorithm
            self._algo_nn = NNAlgorithm(parameters)
        elif algorithm not in ALGORITHMS:
            raise ValueError(f"GRAP: algorithm '{algorithm}' is not implemented")
        if param_space_method not in ('cross', 'pair'):
            raise ValueError("param_space_method must be 'cross' or 'pair'")
        if isinstance(moment_tensors, int):
            moment_tensors = [moment_tensors]
        moment_tensors = sorted(set(int(m) for m in moment_tensors))
        allowed = (0, 1, 2) if legacy_mode else (0, 1, 2, 3, 4, 5)
        if any(m not in allowed for m in moment_tensors):
            # grap.py:434-457 (legacy loop stops at 2), :578-579 (new mode: <= 5)
            raise ValueError("GRAP: moments 0, 1, 2 (legacy mode) / 0 .. 5 (new mode) "
                             "are supported")
        if self._algo_nn is not None:
            self._elements = sorted(list(elements))
            self._algorithm = 'nn'
            self._parameters = self._algo_nn.as_dict()
            self._param_space_method = param_space_method
            self._grid = [{} for _ in range(len(self._algo_nn))]     # one row per filter
            self._moment_tensors = moment_tensors
            self._cutoff_function = cutoff_function
            self._symmetric = symmetric
            self._legacy_mode = legacy_mode
            return
        keys = ALGORITHMS[algorithm]
        if parameters is None:
            parameters = {'eta': [0.05, 4.0, 20.0, 80.0], 'omega': [0.0] * 4} \
                if algorithm == 'sf' else None
        for k in keys:
            if parameters is None or k not in parameters or len(parameters[k]) < 1:
                raise ValueError(f"GRAP/{algorithm}: parameter '{k}' is required")
        self._elements = sorted(list(elements))
        self._algorithm = algorithm
        self._parameters = {k: [float(x) for x in parameters[k]] for k in keys}
        self._param_space_method = param_space_method
        self._grid = _parameter_grid(self._parameters, keys, param_space_method)
        self._moment_tensors = moment_tensors
        self._cutoff_function = cutoff_function
        self._symmetric = symmetric
        self._legacy_mode = legacy_mode

    name = property(lambda self: "GRAP")
    elements = property(lambda self: self._elements)
    cutoff_function = property(lambda self: self._cutoff_function)
    algorithm = property(lambda self: self._algorithm)
    algorithm_object = property(lambda self: self._algo_nn)       # NNAlgorithm or None
    moment_tensors = property(lambda self: self._moment_tensors)
    max_moment = property(lambda self: max(self._moment_tensors))
    is_T_symmetric = property(lambda self: self._symmetric)
    grid = property(lambda self: self._grid)

    def as_dict(self):
        return {"@class": self.__class__.__name__,
                "@module": "tensoralloy.nn.atomic.grap",
                "elements": self._elements, "algorithm": self._algorithm,
                "parameters": self._parameters,
                "param_space_method": self._param_space_method,
                "moment_tensors": self._moment_tensors,
                "cutoff_function": self._cutoff_function,
                "symmetric": self._symmetric, "legacy_mode": self._legacy_mode}

    # -- what the device model consumes ---------------------------------------
    def radial_kind(self):
        return self._algorithm

    def radial_sets(self):
        if self._algo_nn is not None:
            raise ValueError("GRAP/nn has no closed-form parameter sets: the filter network "
                             "is evaluated by nn.atomic.grap_nn.GrapFilterTrainer")
        keys = ALGORITHMS[self._algorithm]
        return [tuple(row[k] for k in keys) for row in self._grid]

    def angular_sets(self):
        return None

    def moments(self):
        """Moments the kernels accumulate, in layout order.  Legacy mode: the requested
        ones (grap.py:419-457); new mode: all of 0..max (grap.py:613, 662-676)."""
        if not self._legacy_mode:
            return tuple(range(self.max_moment + 1))
        return tuple(self._moment_tensors)

    def uses_torch_path(self):
        """True for the descriptors the CUDA descriptor kernels do not serve -- the trainable
        `nn` filter network and moments 4 / 5 (full 3^m moment tensors): they run through
        nn/atomic/grap_nn.py (torch on the device over the library's pair operators)."""
        return self._algorithm == 'nn' or (not self._legacy_mode and self.max_moment > 3)

    def grap_flags(self):
        """include/tab200.h TAB_GRAP_*: the two new-mode details the kernels switch on
        (legacy mode ignores `symmetric`, grap.py:384-466)."""
        if self._legacy_mode:
            return 0
        return 1 | (2 if self._symmetric else 0)

    def dimension(self, angular=False):
        if angular:
            raise ValueError("GRAP is a radial descriptor: use angular=False")
        return len(self._elements) * len(self._grid) * len(self.moments())
