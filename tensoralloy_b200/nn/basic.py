"""
BasicNN: the model base class of the reference (tensoralloy/nn/basic.py:99-1153)
reduced to the energy / force / virial / stress / Hessian outputs of
`BasicNN.build` (:679-787).  TF autograd (`tf.gradients`, :281,:306,:418) is
replaced by the analytic backward kernels of libtab200.
"""
from typing import List

import numpy as np

from tensoralloy_b200.atoms import GPa
from tensoralloy_b200.precision import get_float_dtype
from tensoralloy_b200.utils import Defaults, ModeKeys

API_VERSION = "1.1"


class _PropertyError(ValueError):
    label = "valid"

    def __init__(self, name):
        super().__init__()
        self.name = name

    def __str__(self):
        return f"'{self.name}' is not a '{self.label}' property."


class MinimizablePropertyError(_PropertyError):
    label = "minimizable"


class ExportablePropertyError(_PropertyError):
    label = "exportable"


# basic.py:74-96
_all_properties = (
    ('energy', True, True), ('eentropy', True, True), ('free_energy', True, True),
    ('atomic', False, True), ('forces', True, True), ('stress', True, True),
    ('total_pressure', True, True), ('hessian', False, True),
    ('elastic', True, True), ('rose', True, False), ('eentropy/c', True, False),
    ('hessian/c', True, False), ('extra/c', True, False))
minimizable_properties = [n for n, m, e in _all_properties if m]
exportable_properties = [n for n, m, e in _all_properties if e]

VOIGT = ((0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1))   # basic.py:341-343


class BasicNN:
    default_collection = None
    scope = "Basic"

    def __init__(self, elements: List[str], hidden_sizes=None, activation=None,
                 minimize_properties=('energy', 'forces'),
                 export_properties=('energy', 'forces')):
        self._elements = sorted(list(elements))
        self._hidden_sizes = self._get_hidden_sizes(
            hidden_sizes if hidden_sizes is not None else Defaults.hidden_sizes)
        self._activation = activation or Defaults.activation
        if len(minimize_properties) == 0:
            raise ValueError("At least one property should be minimized.")
        for prop in minimize_properties:
            if prop not in minimizable_properties:
                raise MinimizablePropertyError(prop)
        for prop in export_properties:
            if prop not in exportable_properties:
                raise ExportablePropertyError(prop)
        self._minimize_properties = list(minimize_properties)
        self._export_properties = list(export_properties)
        self._transformer = None

    elements = property(lambda self: self._elements)
    hidden_sizes = property(lambda self: self._hidden_sizes)
    minimize_properties = property(lambda self: self._minimize_properties)
    predict_properties = property(lambda self: self._export_properties)
    transformer = property(lambda self: self._transformer)

    @property
    def is_finite_temperature(self) -> bool:
        return False

    @property
    def variational_energy(self) -> str:
        return "free_energy" if self.is_finite_temperature else "energy"

    def _get_hidden_sizes(self, hidden_sizes):
        results = {}
        for element in self._elements:
            if isinstance(hidden_sizes, dict):
                sizes = np.asarray(hidden_sizes.get(element, Defaults.hidden_sizes),
                                   dtype=int)
            else:
                sizes = np.atleast_1d(hidden_sizes).astype(int)
            assert (sizes > 0).all()
            results[element] = sizes.tolist()
        return results

    def attach_transformer(self, clf):
        self._transformer = clf

    def export(self, output_graph_path: str, checkpoint=None, keep_tmp_files=False,
               use_ema_variables=True, mode=ModeKeys.PREDICT, **kwargs):
        """basic.py:1017-1153.  Writes the frozen `.pb` layout (`Transformer/params`,
        `Metadata/*`, every variable as a Const node under its reference name) that
        `TensorAlloyCalculator(path)` loads; `mode=ModeKeys.NATIVE` writes the
        LAMMPS-native `.npz` instead (basic.py:1037-1039).  The variables are the ones
        this object holds: `checkpoint`, `keep_tmp_files` and `use_ema_variables` are
        accepted for signature parity (no TF checkpoint, no temporary files).  See
        io/graph_model.py `write_graph_model` for what the file is and is not."""
        if self._transformer is None:
            raise ValueError("A transformer must be attached before "
                             "exporting to a pb file.")
        if mode == ModeKeys.NATIVE:
            return self.export_to_lammps_native(output_graph_path, **kwargs)
        from tensoralloy_b200.io.graph_model import write_graph_model
        write_graph_model(self, output_graph_path, precision=kwargs.get('precision'))

    def as_dict(self):
        raise NotImplementedError("This method must be overridden!")

    # ------------------------------------------------------------------
    # evaluation
    # ------------------------------------------------------------------
    def _evaluate(self, features, want_forces, want_virial, want_atomic):
        """Subclass hook: run the kernels.  Returns a dict with float64 numpy
        arrays in LOCAL (caller) atom order: energy (), energy/atom [N],
        forces [N,3], virial [3,3]."""
        raise NotImplementedError

    def build(self, features, mode=ModeKeys.PREDICT, verbose=False):
        """basic.py:679-787.  `features` = transformer.get_constant_features(atoms)
        (the device-side lists).  Returns the reference's prediction dict as
        numpy arrays in the working precision; per-atom arrays are in GSL order
        without the virtual atom, exactly like the TF outputs."""
        if self._transformer is None:
            raise ValueError("A descriptor transformer must be attached.")
        if mode in (ModeKeys.PREDICT, ModeKeys.LAMMPS, ModeKeys.KMC,
                    ModeKeys.PRECOMPUTE, ModeKeys.NATIVE):
            properties = self._export_properties
        else:
            properties = self._minimize_properties
        want_stress = 'stress' in properties or 'total_pressure' in properties \
            or 'elastic' in properties
        want_forces = 'forces' in properties or want_stress
        raw = self._evaluate(features, want_forces, want_stress, True)
        return self._finalize(raw, features, properties)

    # Per-atom results are copied out of the pinned staging buffer into numpy arrays of their
    # own (what a caller of the reference gets from sess.run).  For MD loops over large
    # structures the page faults of a fresh 24 MB array per call cost more than the kernels:
    # with `reuse_result_buffers = True` the arrays of two alternating, persistent sets are
    # overwritten instead -- the results of a call then stay valid until the second next call.
    reuse_result_buffers = False

    def _owned(self, key, src, dtype):
        if not self.reuse_result_buffers:
            return src.astype(dtype)
        pool = self.__dict__.setdefault('_result_pool', [{}, {}])
        flip = self.__dict__.get('_result_flip', 0)
        slot = pool[flip]
        dst = slot.get(key)
        if dst is None or dst.shape != src.shape or dst.dtype != np.dtype(dtype):
            dst = np.empty(src.shape, dtype=dtype)
            slot[key] = dst
        np.copyto(dst, src, casting='unsafe')
        return dst

    def _finalize(self, raw, features, properties):
        dtype = get_float_dtype().as_numpy_dtype
        vap = features.vap
        to_gsl = (lambda a: a) if vap.is_identity else \
            (lambda a: vap.map_array(a.reshape(len(a), -1), reverse=False)[1:].reshape(
                (vap.max_vap_natoms - 1,) + a.shape[1:]))
        self.__dict__['_result_flip'] = 1 - self.__dict__.get('_result_flip', 0)
        pred = {'energy': dtype(raw['energy'])}
        if 'energy/atom' in raw:
            pred['energy/atom'] = self._owned('energy/atom', to_gsl(raw['energy/atom']), dtype)
        if 'forces' in raw:
            pred['forces'] = self._owned('forces', to_gsl(raw['forces']), dtype)
        if 'virial' in raw and ('stress' in properties or
                                'total_pressure' in properties or
                                'elastic' in properties):
            virial = raw['virial']
            stress = virial / features.volume                 # basic.py:317
            voigt = np.array([stress[a, b] for a, b in VOIGT])
            pred['virial'] = virial.astype(dtype)
            pred['stress'] = voigt.astype(dtype)
            # basic.py:393-408: -trace/3 in GPa
            pred['total_pressure'] = dtype(np.trace(stress) / (-3.0 * GPa))
        if 'hessian' in raw:
            pred['hessian'] = raw['hessian'].astype(dtype)
        return pred

    def _eval_flat(self, nbr, n, nb, want_forces, want_virial, want_atomic):
        """Run `model.eval` with every output in ONE device buffer
        [energy nb | virial 9 nb | E_atom n | forces 3 n] and bring it back with a single
        device->host copy (one synchronisation per call instead of one per array)."""
        import torch
        model = self._device_model()
        size = 10 * nb + 4 * n
        buf = getattr(self, '_flat_out', None)
        if buf is None or buf.numel() != size:
            buf = torch.zeros(size, dtype=torch.float64, device='cuda')
            self._flat_out = buf
        o_v, o_e, o_f = nb, 10 * nb, 10 * nb + n
        model.eval(nbr, get_float_dtype().tab_precision, energy=buf[0:nb],
                   eatom=buf[o_e:o_f] if want_atomic else None,
                   forces=buf[o_f:].view(n, 3) if want_forces else None,
                   virial=buf[o_v:o_e] if want_virial else None)
        # one device->host copy into a persistent PINNED buffer; only what was asked for
        hbuf = getattr(self, '_flat_host', None)
        if hbuf is None or hbuf.numel() != size:
            hbuf = torch.zeros(size, dtype=torch.float64).pin_memory()
            self._flat_host = hbuf
        hbuf[:o_e].copy_(buf[:o_e], non_blocking=True)
        if want_atomic:
            hbuf[o_e:o_f].copy_(buf[o_e:o_f], non_blocking=True)
        if want_forces:
            hbuf[o_f:].copy_(buf[o_f:], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        host = hbuf.numpy()
        return (host[0:nb], host[o_v:o_e].reshape(nb, 3, 3), host[o_e:o_f],
                host[o_f:].reshape(n, 3))

    def evaluate_batch(self, batch, want_forces=True, want_virial=True,
                       want_atomic=True):
        """One pass of the kernels over a BATCH of structures
        (`transformer.get_batch_features(images)`): energies [B], per-atom energies and
        forces of all atoms of the batch back to back (caller order, split them with
        `batch.offsets`), virials [B,3,3] -- float64 numpy arrays."""
        if self.is_finite_temperature:
            raise NotImplementedError("batched evaluation of finite-temperature models")
        e, w, ea, f = self._eval_flat(batch.nbr, batch.n_atoms, batch.n_struct,
                                      want_forces, want_virial, want_atomic)
        # (copies: the staging buffer is overwritten by the next call)
        raw = {'energy': e.copy()}
        if want_virial:
            raw['virial'] = w.copy()
        if want_atomic:
            raw['energy/atom'] = ea.copy()
        if want_forces:
            raw['forces'] = f.copy()
        return raw

    def _evaluate_single(self, features, want_forces, want_virial, want_atomic):
        """`_evaluate` of the models whose device handle has an `eval` entry point."""
        e, w, ea, f = self._eval_flat(features.nbr, features.n_atoms, 1, want_forces,
                                      want_virial, want_atomic)
        # views of the pinned staging buffer: `_finalize` copies them into arrays of their own
        raw = {'energy': float(e[0])}
        if want_atomic:
            raw['energy/atom'] = ea
        if want_forces:
            raw['forces'] = f
        if want_virial:
            raw['virial'] = w[0].copy()
        return raw

    def _hessian(self, features):
        """[Nvap, 3, Nvap, 3] Hessian in GSL order incl. the virtual atom (the layout of the
        reference's `Output/Hessian` op, basic.py:410-421).  Models with a closed-form kernel
        (EAM / FS, csrc/hessian.cu) override this; every other model (ADP, AtomicNN, EAM with
        'nn' functions, ...) gets the derivative of its ANALYTIC forces, `_hessian_from_forces`."""
        return self._embed_hessian(self._hessian_from_forces(features), features)

    @staticmethod
    def _embed_hessian(H, features):
        vap = features.vap
        nv = vap.max_vap_natoms
        idx = vap.local_to_gsl_array[1:]
        n = len(idx)
        out = np.zeros((nv, 3, nv, 3), dtype=np.float64)
        out[np.ix_(idx, range(3), idx, range(3))] = np.asarray(H).reshape(n, 3, n, 3)
        return out

    def _hessian_from_forces(self, features, step=1e-4):
        """d2E / dR_a dR_b = -dF_a / dR_b [N,3,N,3] (caller order) from the analytic force
        kernels: fourth-order central differences in every coordinate on the FIXED lists of
        the structure (`tab_nbr_update`), which is what tf.hessians of the reference's graph
        differentiates too (the neighbour metadata are constants of the graph).  The step is
        small (1e-4 A: truncation step^4 F^(5) / 30 is negligible, rounding ~1e-15 / step =
        1e-11 eV/A^2) because cutoff functions are only C1 at rc: a stencil that straddles rc for
        some pair averages a jump of the second derivative, which the pointwise autograd Hessian
        of the reference does not.  tests/test_hessian_gpu.py: 1e-7 eV/A^2 against the oracle's
        autograd Hessian.  4 x 3N force evaluations."""
        import torch
        nbr, d_pos = features.nbr, features.d_pos
        n = features.n_atoms
        base = d_pos.clone()
        H = np.zeros((n, 3, n, 3), dtype=np.float64)
        coef = ((2.0, -1.0), (1.0, 8.0), (-1.0, -8.0), (-2.0, 1.0))
        try:
            for b in range(n):
                for beta in range(3):
                    acc = np.zeros((n, 3))
                    for mult, w in coef:
                        pos = base.clone()
                        pos[b, beta] += mult * step
                        nbr.update(pos)
                        acc += w * self._evaluate(features, True, False, False)['forces']
                    H[:, :, b, beta] = -acc / (12.0 * step)
        finally:
            nbr.update(base)
            torch.cuda.synchronize()
        # the exact Hessian is symmetric: average out the O(step^4) asymmetry
        H2 = H.reshape(3 * n, 3 * n)
        return (0.5 * (H2 + H2.T)).reshape(n, 3, n, 3)

    def _elastic(self, features):
        """Elastic constant tensor [6,6] in GPa with the reference's definition
        (nn/constraint/elastic.py:24-91):
            C_ijkl = [ (d virial_ij / d h)^T h ]_kl / V / GPa
        where the derivative w.r.t. the lattice h is taken at FIXED Cartesian
        positions (only the periodic-image shifts S.h move) -- exact for
        one-atom primitive cells, which is how the reference uses it.  The
        derivative is a central difference of the GPU virial over the 9 lattice
        components (lists and shifts kept, `tab_nbr_update`)."""
        h0 = np.array(features.cell, dtype=np.float64)
        vol = features.volume
        nbr, d_pos = features.nbr, features.d_pos
        step = 1e-5
        dW = np.zeros((3, 3, 3, 3))
        try:
            for m in range(3):
                for k in range(3):
                    w = []
                    for sgn in (+1.0, -1.0):
                        h = h0.copy()
                        h[m, k] += sgn * step
                        nbr.update(d_pos, h)
                        w.append(self._evaluate(features, False, True, False)['virial'])
                    dW[:, :, m, k] = (w[0] - w[1]) / (2.0 * step)
        finally:
            nbr.update(d_pos, h0)
        C = np.einsum('ijmk,ml->ijkl', dW, h0) / vol / GPa
        out = np.zeros((6, 6))
        for vi, (i, j) in enumerate(VOIGT):
            for vj, (k, l) in enumerate(VOIGT):
                out[vi, vj] = C[i, j, k, l]
        return out
