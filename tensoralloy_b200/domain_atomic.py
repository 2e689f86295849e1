"""
Spatial (slab) decomposition of AtomicNN models over the GPUs of one node -- the
AtomicNN row of SURVEY 8(e).  The reference has no spatial decomposition.

Why not the F'(rho)-style second exchange of the EAM path (domain.py): the force on an own
atom i holds dE_j/dR_i of every centre j within rc.  With angular symmetry functions
dE_j/dR_i depends on ALL neighbours k of j (triples j-i-k), so a rank that only knew dE/dG_j
of a halo centre j could still not form its pair gradients without j's whole neighbourhood.
Either the owner of j sends forces back (reverse communication) or the rank recomputes j.
This module recomputes: the halo is 2 rc wide,
    own atoms | inner halo (<= rc from the slab: descriptors, MLP and per-pair gradients are
                recomputed here) | outer halo (rc .. 2 rc: positions only)
and ONE position exchange per step suffices.  The lists are built by `tab_nbr_build_dd` with
[own | inner halo] as the row-owning group; `tab_atomic_eval_dd` masks the inner-halo rows
out of the rank's energy / virial sums; [E, virial] is one 10-double all-reduce.

The same scheme serves the EAM family through `tab_eam_eval_dd` (`EamModel.eval_dd`): it is
the decomposition path of ADP, whose per-species moments would otherwise need their own
exchange, and an alternative to the F' exchange of domain.py for EAM / FS.

`AtomicSlabRank` is comm-agnostic (one rank's device state); `run_loopback` runs all ranks
inside one process on one GPU (test harness); `AtomicSlabDomain` is the torch.distributed
(NCCL send/recv ring) driver.
"""
import numpy as np

from tensoralloy_b200.domain import DistComm, SlabLayout


class AtomicSlabLayout(SlabLayout):
    """Slab geometry with a halo of `halo` = 2 max(rc, acut)."""

    def __init__(self, lx, world, rank, rc_list):
        self.rc_list = float(rc_list)
        self.halo = 2.0 * self.rc_list
        width = float(lx) / int(world)
        if int(world) > 1 and width < self.halo:
            raise ValueError(f"slab width {width:.3f} < 2 rc = {self.halo:.3f}: the halo "
                             f"would span more than the adjacent rank")
        # the base class only needs a cutoff for its own width check
        super().__init__(lx, world, rank, min(self.rc_list, 0.5 * width))

    def send_masks(self, x_owned):
        return x_owned < self.lo + self.halo, x_owned >= self.hi - self.halo

    def frame(self, ly, lz, pad=0.5):
        cell = np.diag([self.width + 2 * self.halo + 2 * pad, ly, lz])
        origin = np.array([self.lo - self.halo - pad, 0.0, 0.0])
        return cell, origin, [0, 1, 1]


class AtomicSlabRank:
    """Device state and kernels of ONE rank."""

    def __init__(self, model, layout, pos_owned, types_owned, ly, lz, precision,
                 device='cuda'):
        import torch
        from tensoralloy_b200 import _lib
        self.torch, self._lib = torch, _lib
        self.model, self.lay, self.precision, self.device = model, layout, precision, device
        self.ly, self.lz = ly, lz
        self.n_own = int(len(pos_owned))
        self.d_pos = torch.as_tensor(np.ascontiguousarray(pos_owned, dtype=np.float64)).to(device)
        self.d_types = torch.as_tensor(np.ascontiguousarray(types_owned, dtype=np.int32)).to(device)
        m_l, m_r = layout.send_masks(np.asarray(pos_owned)[:, 0])
        self.idx_l = torch.from_numpy(np.flatnonzero(m_l)).to(device)
        self.idx_r = torch.from_numpy(np.flatnonzero(m_r)).to(device)
        self.nbr = _lib.NeighborList()
        self.d_out = torch.zeros(16, dtype=torch.float64, device=device)

    def pack(self):
        """(positions, types) of the boundary atoms for the left / right neighbour; the
        sender applies the periodic shift."""
        t = self.torch
        out = []
        for idx, shift in ((self.idx_l, self.lay.shift_to_left),
                           (self.idx_r, self.lay.shift_to_right)):
            p = t.empty((len(idx), 3), dtype=t.float64, device=self.device)
            self._lib.pack_rows(self.d_pos, idx, p, [shift, 0.0, 0.0])
            out.append((p, self.d_types[idx].contiguous()))
        return out

    def receive(self, from_left, from_right):
        """Assemble [own | inner halo | outer halo]."""
        t = self.torch
        pos_h = t.cat([from_left[0], from_right[0]])
        typ_h = t.cat([from_left[1], from_right[1]])
        x = pos_h[:, 0]
        inner = (x >= self.lay.lo - self.lay.rc_list) & (x < self.lay.hi + self.lay.rc_list)
        order = t.cat([t.nonzero(inner).reshape(-1), t.nonzero(~inner).reshape(-1)])
        self.n_inner = int(inner.sum().item())
        self.n_rows = self.n_own + self.n_inner
        self.d_pos_loc = t.cat([self.d_pos, pos_h[order]]).contiguous()
        self.d_types_loc = t.cat([self.d_types, typ_h[order]]).contiguous()
        self.d_mask = t.zeros(self.n_rows, dtype=t.int32, device=self.device)
        self.d_mask[:self.n_own] = 1
        f64 = dict(dtype=t.float64, device=self.device)
        self.d_f = t.zeros((self.n_rows, 3), **f64)
        self.d_ea = t.zeros(self.n_rows, **f64)

    def evaluate(self):
        cell, origin, pbc = self.lay.frame(self.ly, self.lz)
        self.nbr.build_dd(self.d_pos_loc, self.d_types_loc, self.n_rows, cell, origin, pbc,
                          self.lay.rc_list)
        self.model.eval_dd(self.nbr, self.d_mask, self.precision, energy=self.d_out[0:1],
                           eatom=self.d_ea, forces=self.d_f, virial=self.d_out[1:10])

    forces = property(lambda self: self.d_f[:self.n_own])
    eatom = property(lambda self: self.d_ea[:self.n_own])


def _split(pos, types, cell, rc_list, world, model, precision, device):
    pos = np.array(pos, dtype=np.float64)
    lx, ly, lz = cell[0, 0], cell[1, 1], cell[2, 2]
    pos[:, 0] = np.mod(pos[:, 0], lx)
    ranks, owners = [], []
    for r in range(world):
        lay = AtomicSlabLayout(lx, world, r, rc_list)
        mask = lay.owned_mask(pos[:, 0])
        owners.append(np.flatnonzero(mask))
        ranks.append(AtomicSlabRank(model, lay, pos[mask], np.asarray(types)[mask], ly, lz,
                                    precision, device))
    return ranks, owners


def run_loopback(model, pos, types, cell, rc_list, world, precision=0, device='cuda'):
    """Every rank of a `world`-way decomposition inside ONE process on one GPU.
    Returns (E_total, forces [N,3] in input order, virial [3,3], E_atom [N])."""
    import torch
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    ranks, owners = _split(pos, types, cell, rc_list, world, model, precision, device)
    packs = [st.pack() for st in ranks]
    total = torch.zeros(10, dtype=torch.float64, device=device)
    forces = np.zeros((len(pos), 3))
    eatom = np.zeros(len(pos))
    for st, own in zip(ranks, owners):
        # what I receive from my left neighbour is what it sends to ITS right
        st.receive(packs[st.lay.left][1], packs[st.lay.right][0])
        st.evaluate()
        total += st.d_out[:10]
        forces[own] = st.forces.cpu().numpy()
        eatom[own] = st.eatom.cpu().numpy()
    t = total.cpu().numpy()
    return t[0], forces, t[1:10].reshape(3, 3), eatom


class AtomicSlabDomain:
    """One rank of the distributed AtomicNN evaluation (torch.distributed, NCCL ring)."""

    def __init__(self, model, pos, types, cell, rc_list, world, rank, precision=0,
                 device='cuda'):
        import torch
        self.torch = torch
        cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
        pos = np.array(pos, dtype=np.float64)
        pos[:, 0] = np.mod(pos[:, 0], cell[0, 0])
        lay = AtomicSlabLayout(cell[0, 0], world, rank, rc_list)
        mask = lay.owned_mask(pos[:, 0])
        self.owned = np.flatnonzero(mask)
        self.rank_state = AtomicSlabRank(model, lay, pos[mask], np.asarray(types)[mask],
                                         cell[1, 1], cell[2, 2], precision, device)
        self.comm = DistComm(lay) if world > 1 else None
        self.world = world

    def step(self):
        """Exchange the 2 rc halo, evaluate, all-reduce [E, virial].  Returns
        (E_total, own forces [n_own,3] (device), virial [3,3])."""
        t = self.torch
        st = self.rank_state
        (pl, tl), (pr, tr) = st.pack()
        if self.comm is not None:
            dev = st.device
            n_l, n_r = self.comm.exchange_counts(len(pl), len(pr), dev)
            rp_l = t.empty((n_l, 3), dtype=t.float64, device=dev)
            rp_r = t.empty((n_r, 3), dtype=t.float64, device=dev)
            rt_l = t.empty(n_l, dtype=t.int32, device=dev)
            rt_r = t.empty(n_r, dtype=t.int32, device=dev)
            self.comm.exchange(pl, pr, rp_l, rp_r)
            self.comm.exchange(tl, tr, rt_l, rt_r)
            st.receive((rp_l, rt_l), (rp_r, rt_r))
        else:
            st.receive((pr, tr), (pl, tl))      # single rank: its own periodic images
        st.evaluate()
        out = st.d_out[:10].clone()
        if self.comm is not None:
            self.comm.allreduce_sum(out)
        o = out.cpu().numpy()
        return o[0], st.forces, o[1:10].reshape(3, 3)
