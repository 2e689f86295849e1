"""
ctypes binding of libtab200.so (C ABI: include/tab200.h).

There is NO fallback: if the CUDA library is missing or cannot be loaded this
module raises, and so does every product path that needs it.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get('TAB200_LIB', _HERE / 'csrc' / 'libtab200.so'))

TAB_FN_MAX_PARAMS = 32
PRECISION_HIGH = 0
PRECISION_MEDIUM = 1
EAM_ALLOY, EAM_FS, EAM_ADP = 0, 1, 2

FN_ZERO = 0
FN_ZHOU_RHO = 1
FN_ZHOU_PHI = 2
FN_ZHOU_PHI_MIX = 3
FN_ZHOU_EMBED = 4
FN_ZHOU_EMBED_XC = 5
FN_SUTTON_RHO = 6
FN_SUTTON_PHI = 7
FN_SQRT_EMBED = 8
FN_AGRAWAL_RHO = 9
FN_AGRAWAL_PHI = 10
FN_AGRAWAL_EMBED = 11
FN_GRIMES_RHO = 12
FN_GRIMES_PHI = 13
FN_MISHIN_EMBED = 14
FN_MISHIN_POLAR = 15
FN_SPLINE = 16
FN_MLP = 17
FN_POWCUT_RHO = 18
FN_MSAH_PHI = 19
FN_MSAH_EMBED_AL = 20
FN_MSAH_EMBED_FE = 21


class TabFn(C.Structure):
    _fields_ = [('kind', C.c_int32), ('aux', C.c_int32),
                ('p', C.c_double * TAB_FN_MAX_PARAMS)]


def make_fn(kind, params=(), aux=0):
    fn = TabFn()
    fn.kind = int(kind)
    fn.aux = int(aux)
    if len(params) > TAB_FN_MAX_PARAMS:
        raise ValueError("too many parameters for one tab_fn")
    for k, v in enumerate(params):
        fn.p[k] = float(v)
    return fn


CUTOFFS = {'cosine': 0, 'polynomial': 1}
ACTIVATIONS = {'softplus': 0, 'tanh': 1, 'relu': 2, 'leaky_relu': 3, 'sigmoid': 4,
               'softsign': 5, 'elu': 6, 'squareplus': 7}
_DP = C.POINTER(C.c_double)


class TabSfDesc(C.Structure):
    _fields_ = [('n_el', C.c_int32), ('cutoff', C.c_int32), ('angular', C.c_int32),
                ('n_r', C.c_int32), ('n_a', C.c_int32),
                ('rc', C.c_double), ('acut', C.c_double),
                ('eta', _DP), ('omega', _DP), ('beta', _DP), ('gamma', _DP),
                ('zeta', _DP),
                ('radial_kind', C.c_int32), ('n_moments', C.c_int32),
                ('moments', C.c_int32 * 4), ('grap_flags', C.c_int32), ('p3', _DP)]


RADIAL_KINDS = {'sf': 0, 'morse': 1, 'density': 2, 'pexp': 3}
GRAP_SIGNED_SQRT_M0 = 1      # include/tab200.h: TAB_GRAP_*
GRAP_TRACELESS = 2


class TabMlpDesc(C.Structure):
    _fields_ = [('n_layers', C.c_int32), ('sizes', C.c_int32 * 9),
                ('activation', C.c_int32), ('use_resnet_dt', C.c_int32),
                ('output_bias', C.c_int32),
                ('weights', _DP * 8), ('biases', _DP * 8),
                ('xlo', _DP), ('xhi', _DP)]


class TabTdDesc(C.Structure):
    _fields_ = [('n_elements', C.c_int32), ('dim', C.c_int32), ('algo', C.c_int32),
                ('special', C.c_int32), ('H', C.POINTER(TabMlpDesc)),
                ('S', C.POINTER(TabMlpDesc)), ('U', C.POINTER(TabMlpDesc))]


class TabError(RuntimeError):
    """A libtab200 call returned a non-zero status."""


_lib = None

# every symbol include/tab200.h declares (tests check the .so exports them all)
EXPORTS = [
    'tab_version', 'tab_last_error',
    'tab_nbr_create', 'tab_nbr_free', 'tab_nbr_build', 'tab_nbr_build_dd',
    'tab_nbr_update', 'tab_pack_rows', 'tab_peer_put', 'tab_sum_slots',
    'tab_nbr_sizes', 'tab_nbr_counts', 'tab_nbr_export',
    'tab_eam_create', 'tab_eam_set_splines', 'tab_model_free', 'tab_eam_eval', 'tab_eam_pass1',
    'tab_eam_pass2', 'tab_eam_hessian', 'tab_eam_compute_host',
    'tab_atomic_create', 'tab_atomic_free', 'tab_atomic_dim', 'tab_atomic_eval',
    'tab_atomic_descriptors', 'tab_atomic_forces', 'tab_atomic_jvp',
    'tab_launch_count', 'tab_launch_count_reset',
    'tab_nbr_build_batch', 'tab_nbr_batch_size',
    'tab_pairs_export', 'tab_pair_forces', 'tab_pair_jvp', 'tab_atomic_eval_dd', 'tab_eam_eval_dd', 'tab_eam_tabulate',
    'tab_profile_enable', 'tab_profile_read',
    'tab_nbr_set_skin', 'tab_nbr_max_displacement', 'tab_nbr_displacement_device',
    'tab_reduce_slots', 'tab_eam_elastic', 'tab_td_create', 'tab_td_free', 'tab_td_eval', 'tab_td_status',
    'tab_dd_partition', 'tab_dd_send_sets',
]


def lib():
    """Load (once) and return the ctypes handle.  Raises if the library has not
    been built: run `python -c "import __graft_entry__ as g; g.build()"`."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(python -c 'import __graft_entry__ as g; g.build()'). "
            "tensoralloy_b200 has no CPU fallback.")
    L = C.CDLL(str(LIB_PATH), mode=getattr(os, 'RTLD_LAZY', 1))
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pp = C.POINTER(vp)
    L.tab_version.restype = C.c_int
    L.tab_last_error.restype = C.c_char_p
    L.tab_nbr_create.argtypes = [pp]
    L.tab_nbr_free.argtypes = [vp]
    L.tab_nbr_build.argtypes = [vp, i32, vp, vp, C.POINTER(dbl), C.POINTER(i32),
                                dbl, vp]
    L.tab_nbr_build_dd.argtypes = [vp, i32, i32, vp, vp, C.POINTER(dbl),
                                   C.POINTER(dbl), C.POINTER(i32), dbl, vp]
    L.tab_nbr_update.argtypes = [vp, vp, C.POINTER(dbl), vp]
    L.tab_nbr_set_skin.argtypes = [vp, dbl]
    L.tab_nbr_max_displacement.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl), vp]
    L.tab_nbr_build_batch.argtypes = [vp, i32, C.POINTER(i32), vp, vp, C.POINTER(dbl),
                                      C.POINTER(i32), dbl, vp]
    L.tab_nbr_batch_size.argtypes = [vp]
    L.tab_pairs_export.argtypes = [vp, vp, vp, vp, vp]
    L.tab_pair_forces.argtypes = [vp, vp, vp, vp, vp]
    L.tab_pair_jvp.argtypes = [vp, vp, vp, vp, vp]
    L.tab_pack_rows.argtypes = [vp, vp, i32, i32, C.POINTER(dbl), vp, vp]
    L.tab_peer_put.argtypes = [vp, i32, vp, i32, i32, vp]
    L.tab_sum_slots.argtypes = [vp, i32, i32, vp, vp]
    L.tab_reduce_slots.argtypes = [vp, i32, i32, i32, vp, vp]
    L.tab_nbr_displacement_device.argtypes = [vp, vp, vp]
    L.tab_nbr_sizes.argtypes = [vp, C.POINTER(i64), C.POINTER(i32),
                                C.POINTER(i32)]
    L.tab_nbr_counts.argtypes = [vp, vp, vp]
    L.tab_nbr_export.argtypes = [vp, vp, vp, vp, vp]
    L.tab_eam_create.argtypes = [pp, i32, i32, C.POINTER(TabFn),
                                 C.POINTER(TabFn), C.POINTER(TabFn),
                                 C.POINTER(TabFn), C.POINTER(TabFn)]
    L.tab_eam_set_splines.argtypes = [vp, C.POINTER(dbl), i64]
    L.tab_model_free.argtypes = [vp]
    L.tab_eam_eval.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    L.tab_eam_pass1.argtypes = [vp, vp, i32, vp, vp]
    L.tab_eam_pass2.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp]
    L.tab_eam_hessian.argtypes = [vp, vp, vp, vp]
    L.tab_eam_elastic.argtypes = [vp, vp, vp, vp]
    L.tab_eam_compute_host.argtypes = [vp, vp, i32, i32, vp, vp,
                                       C.POINTER(dbl), C.POINTER(i32), dbl, i32,
                                       vp, vp, vp, vp, vp]
    L.tab_atomic_create.argtypes = [pp, C.POINTER(TabSfDesc), C.POINTER(TabMlpDesc)]
    L.tab_atomic_free.argtypes = [vp]
    L.tab_atomic_dim.argtypes = [vp]
    L.tab_atomic_eval.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    L.tab_atomic_descriptors.argtypes = [vp, vp, i32, vp, vp]
    L.tab_atomic_eval_dd.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp]
    L.tab_eam_eval_dd.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp]
    L.tab_eam_tabulate.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
    L.tab_atomic_forces.argtypes = [vp, vp, i32, vp, vp, vp, vp]
    L.tab_atomic_jvp.argtypes = [vp, vp, i32, vp, vp, vp, vp]
    L.tab_dd_partition.argtypes = [vp, i32, i32, dbl, dbl, i32, i32, vp, vp, vp, i32, vp, vp,
                                   vp]
    L.tab_dd_send_sets.argtypes = [vp, i32, dbl, dbl, vp, vp, vp, vp, vp]
    L.tab_td_create.argtypes = [pp, C.POINTER(TabTdDesc)]
    L.tab_td_free.argtypes = [vp]
    L.tab_td_status.argtypes = [vp, C.POINTER(i32), vp]
    L.tab_td_eval.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    L.tab_profile_enable.argtypes = [i32]
    L.tab_profile_read.argtypes = [C.POINTER(dbl), C.POINTER(i32)]
    L.tab_launch_count.restype = i64
    L.tab_launch_count_reset.restype = None
    for name in EXPORTS:
        getattr(L, name)
    _lib = L
    return L


def check(status, what=''):
    if status != 0:
        msg = lib().tab_last_error().decode('utf-8', 'replace')
        raise TabError(f"{what} failed with status {status}: {msg}")


def _cell9(cell):
    arr = np.ascontiguousarray(np.asarray(cell, dtype=np.float64).reshape(9))
    return (C.c_double * 9)(*arr.tolist())


def _pbc3(pbc):
    p = np.asarray(pbc).astype(bool).reshape(-1)
    if p.size == 1:
        p = np.repeat(p, 3)
    return (C.c_int32 * 3)(*[int(x) for x in p])


def _ptr(t):
    """Device (or pinned-host) pointer of a torch tensor / None."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pack_rows(src, idx, dst, shift=None):
    """dst[k] = src[idx[k]] (+ shift): one launch of the library's pack kernel."""
    m = int(idx.shape[0])
    ncol = int(src.shape[1]) if src.dim() == 2 else 1
    sh = (C.c_double * 3)(*[float(x) for x in shift]) if shift is not None else None
    check(lib().tab_pack_rows(_ptr(src), _ptr(idx), m, ncol, sh, _ptr(dst), _stream()),
          'tab_pack_rows')


def peer_put(src, peer_ptrs, slot):
    """src[0..n) -> slot `slot` of every peer buffer (device int64 array of addresses)."""
    check(lib().tab_peer_put(_ptr(src), int(src.numel()), _ptr(peer_ptrs),
                             int(peer_ptrs.numel()), int(slot), _stream()), 'tab_peer_put')


def dd_partition(state, lx, width, world, rank, keep, mail_left, mail_right, mail_cap, counts,
                 work):
    """tab_dd_partition: migration of the rows of `state` [n, ncol] (see include/tab200.h)."""
    n, ncol = int(state.shape[0]), int(state.shape[1])
    check(lib().tab_dd_partition(_ptr(state), n, ncol, float(lx), float(width), int(world),
                                 int(rank), _ptr(keep), _ptr(mail_left), _ptr(mail_right),
                                 int(mail_cap), _ptr(counts), _ptr(work), _stream()),
          'tab_dd_partition')


def dd_send_sets(pos, x_left_below, x_right_from, idx_left, idx_right, counts, work):
    """tab_dd_send_sets: indices of the atoms within reach of the low / high face."""
    check(lib().tab_dd_send_sets(_ptr(pos), int(pos.shape[0]), float(x_left_below),
                                 float(x_right_from), _ptr(idx_left), _ptr(idx_right),
                                 _ptr(counts), _ptr(work), _stream()), 'tab_dd_send_sets')


def sum_slots(slots, n_slots, out, n_sum=None):
    """out[q] = sum over the slots for q < n_sum (default: all), max over the slots beyond."""
    n = int(out.numel())
    check(lib().tab_reduce_slots(_ptr(slots), int(n_slots), n, n if n_sum is None else int(n_sum),
                                 _ptr(out), _stream()), 'tab_reduce_slots')


class NeighborList:
    """Owner of one `tab_nbr` handle (cell list + neighbour lists on the GPU)."""

    def __init__(self):
        self._h = C.c_void_p()
        check(lib().tab_nbr_create(C.byref(self._h)), 'tab_nbr_create')
        self.n = 0
        self.n_struct = 0       # > 0: batch handle (build_batch)

    def __del__(self):
        try:
            if self._h:
                lib().tab_nbr_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def build(self, d_pos, d_types, cell, pbc, rc):
        """d_pos: cuda float64 [n,3] contiguous; d_types: cuda int32 [n] or None."""
        import torch
        assert d_pos.is_cuda and d_pos.dtype == torch.float64 and d_pos.is_contiguous()
        if d_types is not None:
            assert d_types.is_cuda and d_types.dtype == torch.int32
        self.n = int(d_pos.shape[0])
        self.n_struct = 0
        self._disp_pending = None        # fresh lists: no displacement reading is in flight
        check(lib().tab_nbr_build(self._h, self.n, _ptr(d_pos), _ptr(d_types),
                                  _cell9(cell), _pbc3(pbc), float(rc), _stream()),
              'tab_nbr_build')

    def build_dd(self, d_pos, d_types, n_owned, cell, origin, pbc, rc):
        """Domain-decomposed build: d_pos = owned atoms then halo atoms."""
        import torch
        assert d_pos.is_cuda and d_pos.dtype == torch.float64 and d_pos.is_contiguous()
        if d_types is not None:
            assert d_types.is_cuda and d_types.dtype == torch.int32
        n_loc = int(d_pos.shape[0])
        self.n = int(n_owned)
        self.n_struct = 0
        org = (C.c_double * 3)(*[float(x) for x in origin])
        check(lib().tab_nbr_build_dd(self._h, self.n, n_loc - self.n, _ptr(d_pos),
                                     _ptr(d_types), _cell9(cell), org, _pbc3(pbc),
                                     float(rc), _stream()), 'tab_nbr_build_dd')

    def build_batch(self, d_pos, d_types, offsets, cells, pbcs, rc):
        """Lists of a BATCH of structures in one handle (tab_nbr_build_batch).
        d_pos [N,3] / d_types [N]: the structures' atoms back to back; offsets
        [B+1]; cells [B,3,3]; pbcs [B,3]."""
        import torch
        assert d_pos.is_cuda and d_pos.dtype == torch.float64 and d_pos.is_contiguous()
        if d_types is not None:
            assert d_types.is_cuda and d_types.dtype == torch.int32
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        nb = len(offsets) - 1
        cells = np.ascontiguousarray(cells, dtype=np.float64).reshape(nb, 9)
        pbcs = np.ascontiguousarray(np.asarray(pbcs).astype(bool), dtype=np.int32).reshape(nb, 3)
        assert int(offsets[-1]) == int(d_pos.shape[0])
        self.n = int(d_pos.shape[0])
        self.n_struct = nb
        self.offsets = offsets
        check(lib().tab_nbr_build_batch(
            self._h, nb, offsets.ctypes.data_as(C.POINTER(C.c_int32)), _ptr(d_pos),
            _ptr(d_types), cells.ctypes.data_as(C.POINTER(C.c_double)),
            pbcs.ctypes.data_as(C.POINTER(C.c_int32)), float(rc), _stream()),
            'tab_nbr_build_batch')

    def update(self, d_pos, cell=None):
        c = _cell9(cell) if cell is not None else None
        check(lib().tab_nbr_update(self._h, _ptr(d_pos), c, _stream()),
              'tab_nbr_update')

    def set_skin(self, skin):
        """Lists of the NEXT build get the radius rc + skin; the pair kernels mask r >= rc
        (include/tab200.h: MD-valid list reuse)."""
        check(lib().tab_nbr_set_skin(self._h, float(skin)), 'tab_nbr_set_skin')
        self.skin = float(skin)

    def max_displacement(self):
        """(largest |R - R_build| of the last update, skin of the current lists)."""
        d, s = C.c_double(), C.c_double()
        check(lib().tab_nbr_max_displacement(self._h, C.byref(d), C.byref(s), _stream()),
              'tab_nbr_max_displacement')
        return float(d.value), float(s.value)

    def displacement_to(self, d_out):
        """max |R - R_build| of the last update -> d_out[0] (device float64), no sync."""
        check(lib().tab_nbr_displacement_device(self._h, _ptr(d_out), _stream()),
              'tab_nbr_displacement_device')

    def step(self, d_pos, d_types, cell, pbc, rc, max_step=None):
        """One MD step of the lists: refresh the positions; rebuild when an atom has moved more
        than half the skin since the last build (always, for lists without a skin).  Returns
        True when the lists were rebuilt.  The result of the following evaluation equals the
        one on freshly built lists (the reference's semantics, universal.py:58).

        `max_step` (an upper bound of any atom's displacement in ONE step, e.g. dt * |v|_max
        from the integrator) makes the decision non-blocking: the displacement of every refresh
        is read back asynchronously, and the lists are rebuilt as soon as the PREVIOUS reading
        plus `max_step` could exceed skin / 2 -- never later than the blocking rule, at most one
        step earlier -- so the host does not wait for the device between the refresh and the
        evaluation kernels."""
        if self.n != int(d_pos.shape[0]) or getattr(self, 'skin', 0.0) <= 0.0 or \
                self.n_struct > 0:
            self.build(d_pos, d_types, cell, pbc, rc)
            self._disp_pending = None
            return True
        if max_step is None:
            self.update(d_pos)
            disp, skin = self.max_displacement()
            self._disp_pending = None
            if not (2.0 * disp <= skin):
                self.build(d_pos, d_types, cell, pbc, rc)
                return True
            return False
        import torch
        pending = getattr(self, '_disp_pending', None)
        if pending is None:
            last = 0.0                   # fresh lists (or the first pipelined call)
        else:
            pending.synchronize()        # the refresh of the previous step: long finished
            last = float(self._h_disp[0])
        if not (2.0 * (last + float(max_step)) <= self.skin):
            self.build(d_pos, d_types, cell, pbc, rc)
            self._disp_pending = None
            return True
        self.update(d_pos)
        if getattr(self, '_d_disp', None) is None:
            self._d_disp = torch.zeros(1, dtype=torch.float64, device='cuda')
            self._h_disp = torch.zeros(1, dtype=torch.float64).pin_memory()
        self.displacement_to(self._d_disp)
        self._h_disp.copy_(self._d_disp, non_blocking=True)
        self._disp_pending = torch.cuda.Event()
        self._disp_pending.record()
        return False

    def sizes(self):
        nij, nnl, next_ = C.c_int64(), C.c_int32(), C.c_int32()
        check(lib().tab_nbr_sizes(self._h, C.byref(nij), C.byref(nnl),
                                  C.byref(next_)), 'tab_nbr_sizes')
        return int(nij.value), int(nnl.value), int(next_.value)

    def counts(self):
        import torch
        out = torch.empty(self.n, dtype=torch.int32, device='cuda')
        check(lib().tab_nbr_counts(self._h, _ptr(out), _stream()), 'tab_nbr_counts')
        return out

    def export(self):
        """(i, j, S) as cuda int32 tensors, the reference's ilist/jlist/n1."""
        import torch
        nij = self.sizes()[0]
        i = torch.empty(max(nij, 1), dtype=torch.int32, device='cuda')
        j = torch.empty(max(nij, 1), dtype=torch.int32, device='cuda')
        S = torch.empty((max(nij, 1), 3), dtype=torch.int32, device='cuda')
        check(lib().tab_nbr_export(self._h, _ptr(i), _ptr(j), _ptr(S), _stream()),
              'tab_nbr_export')
        return i[:nij], j[:nij], S[:nij]


def _nbr_pairs(self):
    """(i, j, D) of every list entry: cuda int32 [nij] x 2 (caller indices), float64
    [nij,3] pair vectors, rows sorted by i (tab_pairs_export)."""
    import torch
    nij = self.sizes()[0]
    i = torch.empty(max(nij, 1), dtype=torch.int32, device='cuda')
    j = torch.empty(max(nij, 1), dtype=torch.int32, device='cuda')
    D = torch.empty((max(nij, 1), 3), dtype=torch.float64, device='cuda')
    check(lib().tab_pairs_export(self._h, _ptr(i), _ptr(j), _ptr(D), _stream()),
          'tab_pairs_export')
    return i[:nij], j[:nij], D[:nij]


def _nbr_pair_forces(self, g, forces, virial):
    check(lib().tab_pair_forces(self._h, _ptr(g), _ptr(forces), _ptr(virial), _stream()),
          'tab_pair_forces')


def _nbr_pair_jvp(self, u, A, out):
    check(lib().tab_pair_jvp(self._h, _ptr(u), _ptr(A), _ptr(out), _stream()),
          'tab_pair_jvp')


NeighborList.pairs = _nbr_pairs
NeighborList.pair_forces = _nbr_pair_forces
NeighborList.pair_jvp = _nbr_pair_jvp


class EamModel:
    """Owner of one `tab_model` handle for an EAM-family potential."""

    def __init__(self, kind, n_el, rho, phi, embed, dipole=None, quadrupole=None):
        nn = n_el * n_el
        assert len(rho) == nn and len(phi) == nn and len(embed) == n_el

        def arr(fns):
            if fns is None:
                return None
            return (TabFn * len(fns))(*fns)

        self._keep = [arr(rho), arr(phi), arr(embed), arr(dipole), arr(quadrupole)]
        self._h = C.c_void_p()
        check(lib().tab_eam_create(C.byref(self._h), int(kind), int(n_el),
                                   *self._keep), 'tab_eam_create')
        self.n_el = n_el
        self.kind = kind

    def set_splines(self, coeffs):
        """coeffs: float64 array [n_intervals_total, 4] (c0, c1, c2, c3)."""
        a = np.ascontiguousarray(np.asarray(coeffs, dtype=np.float64).reshape(-1))
        check(lib().tab_eam_set_splines(self._h, a.ctypes.data_as(C.POINTER(C.c_double)),
                                        int(a.size)), 'tab_eam_set_splines')

    def __del__(self):
        try:
            if self._h:
                lib().tab_model_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def eval(self, nbr, precision=PRECISION_HIGH, energy=None, eatom=None,
             forces=None, virial=None):
        check(lib().tab_eam_eval(self._h, nbr.handle, int(precision), _ptr(energy),
                                 _ptr(eatom), _ptr(forces), _ptr(virial),
                                 _stream()), 'tab_eam_eval')

    def eval_dd(self, nbr, mask, precision=PRECISION_HIGH, energy=None, eatom=None,
                forces=None, virial=None):
        """Decomposed evaluation with recomputed inner-halo rows (tab_eam_eval_dd)."""
        check(lib().tab_eam_eval_dd(self._h, nbr.handle, int(precision), _ptr(mask),
                                    _ptr(energy), _ptr(eatom), _ptr(forces), _ptr(virial),
                                    _stream()), 'tab_eam_eval_dd')

    TABLES = {'rho': 0, 'phi': 1, 'embed': 2, 'dipole': 3, 'quadrupole': 4}

    def tabulate(self, which, index, x):
        """Values and derivatives of one function of the model on the grid `x` (numpy
        float64): tab_eam_tabulate."""
        import torch
        d_x = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to('cuda')
        d_y = torch.empty_like(d_x)
        d_dy = torch.empty_like(d_x)
        check(lib().tab_eam_tabulate(self._h, self.TABLES[which], int(index), int(d_x.numel()),
                                     _ptr(d_x), _ptr(d_y), _ptr(d_dy), _stream()),
              'tab_eam_tabulate')
        return d_y.cpu().numpy(), d_dy.cpu().numpy()

    def pass1(self, nbr, precision, fprime=None):
        check(lib().tab_eam_pass1(self._h, nbr.handle, int(precision), _ptr(fprime),
                                  _stream()), 'tab_eam_pass1')

    def pass2(self, nbr, precision, fprime_halo=None, energy=None, eatom=None,
              forces=None, virial=None):
        check(lib().tab_eam_pass2(self._h, nbr.handle, int(precision),
                                  _ptr(fprime_halo), _ptr(energy), _ptr(eatom),
                                  _ptr(forces), _ptr(virial), _stream()),
              'tab_eam_pass2')

    def elastic(self, nbr):
        """[6, 6] = sum over the atoms of tab_eam_elastic's shares: (d virial / d h)^T h in
        Voigt pairs, eV (divide by V GPa for the reference's elastic tensor)."""
        import torch
        out = torch.empty((nbr.n, 36), dtype=torch.float64, device='cuda')
        check(lib().tab_eam_elastic(self._h, nbr.handle, _ptr(out), _stream()),
              'tab_eam_elastic')
        return out.sum(dim=0).reshape(6, 6)

    def hessian(self, nbr):
        """Dense [n,3,n,3] float64 Hessian (caller atom order) on the device."""
        import torch
        n = nbr.n
        H = torch.empty((n, 3, n, 3), dtype=torch.float64, device='cuda')
        check(lib().tab_eam_hessian(self._h, nbr.handle, _ptr(H), _stream()),
              'tab_eam_hessian')
        return H

    def compute_host(self, nbr, precision, h_pos, h_types, cell, pbc, rc, rebuild,
                     h_energy, h_eatom, h_forces, h_virial):
        """All arguments are HOST torch tensors (pinned for full speed)."""
        n = int(h_pos.shape[0])
        check(lib().tab_eam_compute_host(
            self._h, nbr.handle, int(precision), n, _ptr(h_pos), _ptr(h_types),
            _cell9(cell), _pbc3(pbc), float(rc), int(bool(rebuild)),
            _ptr(h_energy), _ptr(h_eatom), _ptr(h_forces), _ptr(h_virial),
            _stream()), 'tab_eam_compute_host')
        nbr.n = n


def profile_enable(on=True):
    check(lib().tab_profile_enable(int(bool(on))), 'tab_profile_enable')


def profile_read():
    """Mean milliseconds of (rho pass, F' spread, force pass, reduction), calls."""
    ms = (C.c_double * 4)()
    calls = C.c_int32()
    check(lib().tab_profile_read(ms, C.byref(calls)), 'tab_profile_read')
    return list(ms), int(calls.value)


def _fill_mlp_desc(d, m, keep):
    """One network -> TabMlpDesc.  m: dict(weights=[np [in, out] ...], biases=[np or None ...],
    activation, use_resnet_dt, output_bias, xlo, xhi); `keep` collects the arrays whose
    memory the descriptor points to."""
    def darr(vals):
        a = np.ascontiguousarray(np.asarray(vals, dtype=np.float64).reshape(-1))
        keep.append(a)
        return a.ctypes.data_as(_DP)

    W = m['weights']
    if len(W) > 8:
        raise ValueError("at most 8 layers (hidden + output) are supported")
    d.n_layers = len(W)
    for k, w in enumerate(W):
        w = np.asarray(w, dtype=np.float64)
        d.sizes[k] = w.shape[0]
        d.sizes[k + 1] = w.shape[1]
        d.weights[k] = darr(w)
        b = m['biases'][k] if k < len(m['biases']) else None
        d.biases[k] = darr(b) if b is not None else None
    d.activation = ACTIVATIONS[m.get('activation', 'softplus').lower()]
    d.use_resnet_dt = int(bool(m.get('use_resnet_dt', False)))
    d.output_bias = int(bool(m.get('output_bias', False)))
    if m.get('xlo') is not None and m.get('xhi') is not None:
        d.xlo = darr(m['xlo'])
        d.xhi = darr(m['xhi'])


class TdHeads:
    """Owner of one `tab_td` handle: the H / S / U networks of the finite-temperature model
    (`tab_td_eval`: U, S, F per atom and dF/dG in one kernel).

    heads : per element (sorted order) dict(H=..., S=..., U=...) of network dicts as
            `_fill_mlp_desc` takes them (H carries xlo / xhi)
    algo  : 'default' | 'Sommerfeld';  special : None | 'Be'
    """

    def __init__(self, dim, heads, algo='default', special=None):
        self._keep = []
        n_el = len(heads)
        self._H = (TabMlpDesc * n_el)()
        self._S = (TabMlpDesc * n_el)()
        self._U = (TabMlpDesc * n_el)()
        for e, h in enumerate(heads):
            _fill_mlp_desc(self._H[e], h['H'], self._keep)
            _fill_mlp_desc(self._S[e], h['S'], self._keep)
            _fill_mlp_desc(self._U[e], h['U'], self._keep)
        d = TabTdDesc()
        d.n_elements, d.dim = n_el, int(dim)
        d.algo = {'default': 0, 'sommerfeld': 1}[str(algo).lower()]
        d.special = {None: 0, 'Be': 1}[special]
        d.H, d.S, d.U = self._H, self._S, self._U
        self.dim = int(dim)
        self._h = C.c_void_p()
        check(lib().tab_td_create(C.byref(self._h), C.byref(d)), 'tab_td_create')

    def __del__(self):
        try:
            if self._h:
                lib().tab_td_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def eval(self, types, G, T, precision=PRECISION_HIGH):
        """types int32 [n], G float64 [n, dim], T float64 [n] (cuda) ->
        (U, S, F [n], dFdG [n, dim]) float64 cuda tensors."""
        import torch
        n = int(types.shape[0])
        out = torch.empty((3, n), dtype=torch.float64, device='cuda')
        dfdg = torch.empty((n, self.dim), dtype=torch.float64, device='cuda')
        check(lib().tab_td_eval(self._h, n, _ptr(types), _ptr(G), _ptr(T), int(precision),
                                _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(dfdg),
                                _stream()), 'tab_td_eval')
        return out[0], out[1], out[2], dfdg

    def status(self):
        """0, or 1 when a weight transfer of an earlier `eval` timed out (outputs are NaN)."""
        flag = C.c_int32(0)
        check(lib().tab_td_status(self._h, C.byref(flag), _stream()), 'tab_td_status')
        return int(flag.value)


class AtomicModel:
    """Owner of one `tab_atomic` handle (symmetry functions + per-element MLPs).

    radial  : list of (eta, omega);  angular : list of (beta, gamma, zeta) or None
    mlps    : per element (sorted order) dict(weights=[np [in,out]...],
              biases=[np [out] or None...], activation=str, use_resnet_dt=bool,
              output_bias=bool, xlo=np or None, xhi=np or None)
    """

    def __init__(self, n_el, rc, acut, radial, angular, cutoff, mlps,
                 radial_kind='sf', moments=(0,), grap_flags=0):
        """radial: list of parameter tuples (2 or 3 values per set, see
        include/tab200.h: tab_sf_desc); moments: GRAP multipole moments; grap_flags:
        GRAP_SIGNED_SQRT_M0 | GRAP_TRACELESS (new mode of the reference)."""
        self._keep = []

        def darr(vals):
            a = np.ascontiguousarray(np.asarray(vals, dtype=np.float64).reshape(-1))
            self._keep.append(a)
            return a.ctypes.data_as(_DP)

        sf = TabSfDesc()
        sf.n_el = n_el
        sf.cutoff = CUTOFFS[cutoff]
        sf.angular = 1 if angular else 0
        sf.n_r = len(radial)
        sf.n_a = len(angular) if angular else 0
        sf.rc = float(rc)
        sf.acut = float(acut if acut is not None else rc)
        sf.eta = darr([r[0] for r in radial])
        sf.omega = darr([r[1] for r in radial])
        sf.p3 = darr([r[2] if len(r) > 2 else 0.0 for r in radial])
        sf.radial_kind = RADIAL_KINDS[radial_kind]
        moments = sorted(set(int(x) for x in moments))
        if len(moments) > 4:
            raise ValueError("GRAP moments: subset of 0..3")
        sf.n_moments = len(moments)
        sf.grap_flags = int(grap_flags)
        for k, mm in enumerate(moments):
            sf.moments[k] = mm
        ang = angular or []
        sf.beta = darr([a[0] for a in ang] or [0.0])
        sf.gamma = darr([a[1] for a in ang] or [0.0])
        sf.zeta = darr([a[2] for a in ang] or [0.0])
        descs = (TabMlpDesc * n_el)()
        for e, m in enumerate(mlps):
            d = descs[e]
            W = m['weights']
            d.n_layers = len(W)
            if len(W) > 8:
                raise ValueError("at most 8 layers (hidden + output) are supported")
            for k, w in enumerate(W):
                w = np.asarray(w, dtype=np.float64)
                d.sizes[k] = w.shape[0]
                d.sizes[k + 1] = w.shape[1]
                d.weights[k] = darr(w)
                b = m['biases'][k] if k < len(m['biases']) else None
                d.biases[k] = darr(b) if b is not None else None
            d.activation = ACTIVATIONS[m.get('activation', 'softplus').lower()]
            d.use_resnet_dt = int(bool(m.get('use_resnet_dt', False)))
            d.output_bias = int(bool(m.get('output_bias', False)))
            if m.get('xlo') is not None and m.get('xhi') is not None:
                d.xlo = darr(m['xlo'])
                d.xhi = darr(m['xhi'])
        self._h = C.c_void_p()
        check(lib().tab_atomic_create(C.byref(self._h), C.byref(sf), descs),
              'tab_atomic_create')
        self.dim = int(lib().tab_atomic_dim(self._h))

    def __del__(self):
        try:
            if self._h:
                lib().tab_atomic_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def eval(self, nbr, precision=PRECISION_HIGH, energy=None, eatom=None,
             forces=None, virial=None):
        check(lib().tab_atomic_eval(self._h, nbr.handle, int(precision), _ptr(energy),
                                    _ptr(eatom), _ptr(forces), _ptr(virial),
                                    _stream()), 'tab_atomic_eval')

    def eval_dd(self, nbr, mask, precision=PRECISION_HIGH, energy=None, eatom=None,
                forces=None, virial=None):
        """Spatial decomposition: lists built with build_dd over [own | inner halo] +
        outer halo; `mask` (cuda int32) marks the own atoms (tab_atomic_eval_dd)."""
        check(lib().tab_atomic_eval_dd(self._h, nbr.handle, int(precision), _ptr(mask),
                                       _ptr(energy), _ptr(eatom), _ptr(forces),
                                       _ptr(virial), _stream()), 'tab_atomic_eval_dd')

    def forces_from_dedg(self, nbr, dedg, forces, virial, precision=PRECISION_HIGH):
        check(lib().tab_atomic_forces(self._h, nbr.handle, int(precision), _ptr(dedg),
                                      _ptr(forces), _ptr(virial), _stream()),
              'tab_atomic_forces')

    def jvp(self, nbr, u, A, out, precision=PRECISION_HIGH):
        check(lib().tab_atomic_jvp(self._h, nbr.handle, int(precision), _ptr(u), _ptr(A),
                                   _ptr(out), _stream()), 'tab_atomic_jvp')

    def descriptors(self, nbr, precision=PRECISION_HIGH):
        import torch
        out = torch.zeros((nbr.n, self.dim), dtype=torch.float64, device='cuda')
        check(lib().tab_atomic_descriptors(self._h, nbr.handle, int(precision),
                                           _ptr(out), _stream()),
              'tab_atomic_descriptors')
        return out
