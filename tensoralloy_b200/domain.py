"""
Spatial (1-D slab) decomposition of the EAM force step over the GPUs of one box.

The reference has no spatial decomposition; this is the new
multi-GPU path of SURVEY.md 8(e):

  rank r owns the atoms with  lo_r <= x < hi_r  (slabs along x, periodic ring);
  per step
    1. pack the positions of the atoms within rc of the two slab faces, exchange
       them with the two ring neighbours (halo = rc; atoms crossing the periodic
       boundary are shifted by -/+ Lx by the sender)
    2. refresh (or rebuild) the local lists over owned + halo atoms
       (libtab200: tab_nbr_update / tab_nbr_build_dd; y and z stay periodic
       inside the library)
    3. pass 1 on the owned atoms: rho_i, F(rho_i), F'(rho_i)
    4. exchange F' of the same boundary atoms (8 B per halo atom)
    5. pass 2 on the owned atoms: forces, energies, partial virial
    6. all-reduce of [E, virial(9)]  (10 doubles)
  With FULL neighbour lists an owned atom never needs a ghost's force, only its
  position and F'(rho): no reverse communication.

`DistComm` drives the exchange over torch.distributed (NCCL on GPUs, gloo in the
CPU tests); `run_loopback` runs every rank inside one process to test the whole
pipeline on a single GPU.
"""
import numpy as np


class SlabLayout:
    """Pure geometry: which atoms a rank owns and which it sends where."""

    def __init__(self, lx, world, rank, rc):
        self.lx = float(lx)
        self.world = int(world)
        self.rank = int(rank)
        self.rc = float(rc)
        self.width = self.lx / self.world
        if self.world > 1 and self.width < 2.0 * rc:
            raise ValueError(
                f"slab width {self.width:.3f} < 2 rc: the halo would span more "
                f"than the adjacent rank")
        self.lo = self.rank * self.width
        self.hi = (self.rank + 1) * self.width
        self.left = (self.rank - 1) % self.world
        self.right = (self.rank + 1) % self.world
        # shift applied by the SENDER so that the receiver sees contiguous space
        self.shift_to_left = self.lx if self.rank == 0 else 0.0
        self.shift_to_right = -self.lx if self.rank == self.world - 1 else 0.0

    def owned_mask(self, x_wrapped):
        if self.rank == self.world - 1:
            return (x_wrapped >= self.lo)
        return (x_wrapped >= self.lo) & (x_wrapped < self.hi)

    def send_masks(self, x_owned):
        """Atoms within rc of the low / high face (x_owned already wrapped)."""
        return x_owned < self.lo + self.rc, x_owned >= self.hi - self.rc

    def frame(self, ly, lz, pad=0.5):
        """Binning frame of the local system handed to tab_nbr_build_dd."""
        cell = np.diag([self.width + 2 * self.rc + 2 * pad, ly, lz])
        origin = np.array([self.lo - self.rc - pad, 0.0, 0.0])
        return cell, origin, [0, 1, 1]


class DistComm:
    """Ring exchange + all-reduce over torch.distributed."""

    def __init__(self, layout):
        import torch.distributed as dist
        self.dist = dist
        self.layout = layout

    def exchange(self, send_left, send_right, recv_from_left, recv_from_right):
        dist = self.dist
        lay = self.layout
        # order matters when left == right (world == 2): the peer's "to_left"
        # message is our "from_right" one
        ops = [dist.P2POp(dist.isend, send_left, lay.left),
               dist.P2POp(dist.isend, send_right, lay.right),
               dist.P2POp(dist.irecv, recv_from_right, lay.right),
               dist.P2POp(dist.irecv, recv_from_left, lay.left)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def exchange_counts(self, n_left, n_right, device):
        import torch
        s_l = torch.tensor([n_left], dtype=torch.int64, device=device)
        s_r = torch.tensor([n_right], dtype=torch.int64, device=device)
        r_l = torch.zeros(1, dtype=torch.int64, device=device)
        r_r = torch.zeros(1, dtype=torch.int64, device=device)
        self.exchange(s_l, s_r, r_l, r_r)
        return int(r_l.item()), int(r_r.item())

    def allreduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)


class SlabRank:
    """Device state and kernels of ONE rank (comm-agnostic)."""

    def __init__(self, model, layout, pos_owned, ly, lz, precision, device='cuda'):
        import torch
        from tensoralloy_b200 import _lib
        self.torch = torch
        self._lib = _lib
        self.model = model
        self.lay = layout
        self.precision = precision
        self.device = device
        self.ly, self.lz = ly, lz
        self.n_owned = int(len(pos_owned))
        self.h_pos = torch.from_numpy(np.ascontiguousarray(pos_owned)).pin_memory()
        self._pos0 = self.h_pos.to(device)
        m_l, m_r = layout.send_masks(pos_owned[:, 0])
        self.idx_l = torch.from_numpy(np.flatnonzero(m_l)).to(device)
        self.idx_r = torch.from_numpy(np.flatnonzero(m_r)).to(device)
        self.shift_l = [layout.shift_to_left, 0.0, 0.0]
        self.shift_r = [layout.shift_to_right, 0.0, 0.0]
        f64 = dict(dtype=torch.float64, device=device)
        self.send_pos_l = torch.empty((len(self.idx_l), 3), **f64)
        self.send_pos_r = torch.empty((len(self.idx_r), 3), **f64)
        self.send_fp_l = torch.empty(len(self.idx_l), **f64)
        self.send_fp_r = torch.empty(len(self.idx_r), **f64)
        self.nbr = _lib.NeighborList()
        self.n_from_l = self.n_from_r = 0
        self.d_pos_loc = None
        self.d_fp = torch.zeros(self.n_owned, dtype=torch.float64, device=device)
        self.d_out = torch.zeros(16, dtype=torch.float64, device=device)
        self.d_f = torch.zeros((self.n_owned, 3), dtype=torch.float64, device=device)
        self.h_out = torch.zeros(16, dtype=torch.float64).pin_memory()
        self.h_f = torch.zeros((self.n_owned, 3), dtype=torch.float64).pin_memory()

    # -- halo bookkeeping ----------------------------------------------------
    def set_halo_counts(self, n_from_left, n_from_right):
        t = self.torch
        self.n_from_l, self.n_from_r = n_from_left, n_from_right
        n_halo = n_from_left + n_from_right
        self.d_pos_loc = t.empty((self.n_owned + n_halo, 3), dtype=t.float64,
                                 device=self.device)
        # the owned positions LIVE in the head of the local array (no copy per step)
        self.d_pos_owned = self.d_pos_loc[:self.n_owned]
        self.d_pos_owned.copy_(self._pos0)
        self.d_fp_halo = t.zeros(max(n_halo, 1), dtype=t.float64, device=self.device)
        o = self.n_owned
        self.recv_pos_l = self.d_pos_loc[o:o + n_from_left]
        self.recv_pos_r = self.d_pos_loc[o + n_from_left:]
        self.recv_fp_l = self.d_fp_halo[:n_from_left]
        self.recv_fp_r = self.d_fp_halo[n_from_left:n_halo]

    def pack_positions(self):
        p = self.d_pos_owned
        self._lib.pack_rows(p, self.idx_l, self.send_pos_l, self.shift_l)
        self._lib.pack_rows(p, self.idx_r, self.send_pos_r, self.shift_r)
        return self.send_pos_l, self.send_pos_r

    def pack_fprime(self):
        self._lib.pack_rows(self.d_fp, self.idx_l, self.send_fp_l)
        self._lib.pack_rows(self.d_fp, self.idx_r, self.send_fp_r)
        return self.send_fp_l, self.send_fp_r

    # -- kernels -------------------------------------------------------------
    def build(self):
        cell, origin, pbc = self.lay.frame(self.ly, self.lz)
        self.nbr.build_dd(self.d_pos_loc, None, self.n_owned, cell, origin, pbc,
                          self.lay.rc)

    def update(self):
        self.nbr.update(self.d_pos_loc)

    def pass1(self):
        self.model.pass1(self.nbr, self.precision, fprime=self.d_fp)

    def pass2(self):
        self.model.pass2(self.nbr, self.precision,
                         fprime_halo=self.d_fp_halo if self.n_from_l + self.n_from_r
                         else None,
                         energy=self.d_out[0:1], forces=self.d_f,
                         virial=self.d_out[1:10])


class SlabDomain:
    """One rank of the distributed MD force step (used by bench.py)."""

    def __init__(self, model, cells, a, rc, sigma, seed, world, rank,
                 scaling='strong', precision=0, device='cuda'):
        from tensoralloy_b200.atoms import fcc_positions
        gx = cells * world if scaling == 'weak' else cells
        pos, cell = fcc_positions(a, gx, cells, cells)
        rng = np.random.default_rng(seed)
        pos = pos + rng.normal(scale=sigma, size=pos.shape)
        lx, ly, lz = cell[0, 0], cell[1, 1], cell[2, 2]
        pos[:, 0] = np.mod(pos[:, 0], lx)
        self.n_total = len(pos)
        self.layout = SlabLayout(lx, world, rank, rc)
        owned = pos[self.layout.owned_mask(pos[:, 0])]
        del pos
        self.rank_state = SlabRank(model, self.layout, owned, ly, lz, precision, device)
        self.comm = DistComm(self.layout)
        self.scaling = scaling
        r = self.rank_state
        n_l, n_r = self.comm.exchange_counts(len(r.idx_l), len(r.idx_r), device)
        r.set_halo_counts(n_l, n_r)
        self._exchange_positions()
        r.build()
        self.n_local = r.n_owned
        self.nij_local = r.nbr.sizes()[0]
        self.h2d_bytes = r.n_owned * 24
        self.d2h_bytes = r.n_owned * 24 + 80

    def describe(self):
        lay = self.layout
        return (f"1-D slabs along x, {lay.world} ranks x {self.n_local} owned atoms "
                f"(this rank), halo = rc = {lay.rc} A, 2 NCCL ring exchanges "
                f"(positions, F') + one 10-double all-reduce per step; "
                f"{self.scaling} scaling")

    def _exchange_positions(self):
        r = self.rank_state
        s_l, s_r = r.pack_positions()
        self.comm.exchange(s_l, s_r, r.recv_pos_l, r.recv_pos_r)

    def _exchange_fprime(self):
        r = self.rank_state
        s_l, s_r = r.pack_fprime()
        self.comm.exchange(s_l, s_r, r.recv_fp_l, r.recv_fp_r)

    def step(self):
        r = self.rank_state
        self._exchange_positions()
        r.update()
        r.pass1()
        self._exchange_fprime()
        r.pass2()
        self.comm.allreduce_sum(r.d_out[:10])

    def step_e2e(self):
        """Host positions in, host forces / energy / virial out, lists rebuilt."""
        r = self.rank_state
        r.d_pos_owned.copy_(r.h_pos, non_blocking=True)
        self._exchange_positions()
        r.build()
        r.pass1()
        self._exchange_fprime()
        r.pass2()
        self.comm.allreduce_sum(r.d_out[:10])
        r.h_f.copy_(r.d_f, non_blocking=True)
        r.h_out.copy_(r.d_out, non_blocking=True)
        r.torch.cuda.synchronize()

    def results(self):
        r = self.rank_state
        out = r.d_out.cpu().numpy()
        return out[0], r.d_f.cpu().numpy(), out[1:10].reshape(3, 3)


def run_loopback(model, pos, cell, rc, world, precision=0, rebuild=True, device='cuda'):
    """Run every rank of a `world`-way slab decomposition inside ONE process on
    one GPU (test harness for the decomposition logic + the DD kernels).
    Returns (E_total, forces[N,3] in input order, virial[3,3])."""
    import torch
    pos = np.array(pos, dtype=np.float64)
    lx, ly, lz = cell[0, 0], cell[1, 1], cell[2, 2]
    pos[:, 0] = np.mod(pos[:, 0], lx)
    ranks, owners = [], []
    for r in range(world):
        lay = SlabLayout(lx, world, r, rc)
        mask = lay.owned_mask(pos[:, 0])
        owners.append(np.flatnonzero(mask))
        ranks.append(SlabRank(model, lay, pos[mask], ly, lz, precision, device))
    for r, st in enumerate(ranks):
        left, right = ranks[st.lay.left], ranks[st.lay.right]
        # what I receive from my left neighbour is what it sends to ITS right
        st.set_halo_counts(len(left.idx_r), len(right.idx_l))

    def exchange(kind):
        packs = [getattr(st, kind)() for st in ranks]
        for st in ranks:
            from_left = packs[st.lay.left][1]
            from_right = packs[st.lay.right][0]
            if kind == 'pack_positions':
                st.recv_pos_l.copy_(from_left)
                st.recv_pos_r.copy_(from_right)
            else:
                st.recv_fp_l.copy_(from_left)
                st.recv_fp_r.copy_(from_right)

    exchange('pack_positions')
    for st in ranks:
        st.build() if rebuild else st.update()
        st.pass1()
    exchange('pack_fprime')
    total = torch.zeros(10, dtype=torch.float64, device=device)
    forces = np.zeros_like(pos)
    for st, own in zip(ranks, owners):
        st.pass2()
        total += st.d_out[:10]
        forces[own] = st.d_f.cpu().numpy()
    t = total.cpu().numpy()
    return t[0], forces, t[1:10].reshape(3, 3)
