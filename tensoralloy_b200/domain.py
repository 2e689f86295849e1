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

`PeerComm` is the NVLink path: every rank's local position / F' arrays live in
torch symmetric memory, the pack kernels of a rank store STRAIGHT into its ring
neighbours' receive regions over NVLink (pack + send are one kernel, no staging
buffer, no NCCL call) and a device-side barrier orders the steps; the 10-double
reduction is a one-shot all-reduce over the same peer mappings.  `DistComm` is
the fallback over torch.distributed point-to-point (NCCL on GPUs, gloo in the CPU
tests).  The resident-list step is captured once in a CUDA graph (`enable_graph`),
so a step is ONE launch per rank.  `run_loopback` runs every rank inside one
process to test the whole pipeline on a single GPU.
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


class PeerComm:
    """Halo exchange by direct stores into the neighbours' memory (NVLink / NVSwitch
    peer mappings from torch symmetric memory).

    One flat float64 symmetric buffer per rank, same capacity everywhere:
        [ positions of owned | from-left | from-right atoms  (3 doubles each) ]
        [ F' of from-left | from-right atoms ]
    and a 16-double symmetric buffer for [E, virial(9)].  `dst_*` are views of the
    NEIGHBOURS' buffers: the region of the left neighbour that holds what it
    receives from its right (= me), and vice versa.
    """

    def __init__(self, layout, n_owned, n_send_left, n_send_right, device):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.torch, self.dist, self.layout = torch, dist, layout
        world, rank = layout.world, layout.rank
        # what each rank owns / receives: from-left of r = send-right of r-1, ...
        mine = torch.tensor([n_owned, n_send_left, n_send_right], dtype=torch.int64,
                            device=device)
        table = [torch.zeros(3, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(table, mine)
        table = [t.tolist() for t in table]
        owned = [t[0] for t in table]
        from_l = [table[(r - 1) % world][2] for r in range(world)]
        from_r = [table[(r + 1) % world][1] for r in range(world)]
        rows_cap = max(owned[r] + from_l[r] + from_r[r] for r in range(world))
        halo_cap = max(from_l[r] + from_r[r] for r in range(world))
        self.n_from_l, self.n_from_r = from_l[rank], from_r[rank]
        self.off_fp = 3 * rows_cap
        total = self.off_fp + max(halo_cap, 1)
        group = dist.group.WORLD
        self.group_name = group.group_name
        self.buf = symm.empty(total, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        # all-reduce of [E, virial]: slot r of every rank's `red_all` is written by rank r
        self.red_all = symm.empty(16 * world, dtype=torch.float64, device=device)
        self.red_all.zero_()
        self.red_hdl = symm.rendezvous(self.red_all, group)
        self.red_ptrs = torch.tensor([int(p) for p in self.red_hdl.buffer_ptrs],
                                     dtype=torch.int64, device=device)
        self.red = torch.zeros(16, dtype=torch.float64, device=device)      # my partial sums
        self.red_out = torch.zeros(16, dtype=torch.float64, device=device)
        lay = layout
        L, R = lay.left, lay.right
        f64 = torch.float64
        # my views
        rows = owned[rank] + self.n_from_l + self.n_from_r
        self.pos_loc = self.buf[:3 * rows].view(rows, 3)
        n_halo = self.n_from_l + self.n_from_r
        self.fp_halo = self.buf[self.off_fp:self.off_fp + max(n_halo, 1)]
        # neighbours' receive regions
        self.dst_pos_l = self.hdl.get_buffer(L, (n_send_left, 3), f64,
                                             3 * (owned[L] + from_l[L]))
        self.dst_pos_r = self.hdl.get_buffer(R, (n_send_right, 3), f64, 3 * owned[R])
        self.dst_fp_l = self.hdl.get_buffer(L, (n_send_left,), f64, self.off_fp + from_l[L])
        self.dst_fp_r = self.hdl.get_buffer(R, (n_send_right,), f64, self.off_fp)
        torch.cuda.synchronize()
        dist.barrier()

    def fence(self, channel):
        """All ranks have issued (and completed) the stores before this point."""
        self.hdl.barrier(channel=channel)

    def allreduce_sum(self):
        """self.red summed over the ranks -> self.red_out: one kernel stores my 16
        doubles into my slot of EVERY rank's buffer over NVLink, a device barrier,
        one kernel sums the slots in rank order (deterministic, identical on all
        ranks).  (torch's one-shot symmetric all-reduce has no float64 kernel.)"""
        from tensoralloy_b200 import _lib
        _lib.peer_put(self.red, self.red_ptrs, self.layout.rank)
        self.red_hdl.barrier(channel=2)
        _lib.sum_slots(self.red_all, self.layout.world, self.red_out)
        return self.red_out


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
    def set_halo_counts(self, n_from_left, n_from_right, pos_loc=None, fp_halo=None,
                        out=None):
        """`pos_loc` / `fp_halo` / `out`: externally owned storage (the symmetric
        buffers of PeerComm) instead of private allocations."""
        t = self.torch
        self.n_from_l, self.n_from_r = n_from_left, n_from_right
        n_halo = n_from_left + n_from_right
        self.d_pos_loc = pos_loc if pos_loc is not None else \
            t.empty((self.n_owned + n_halo, 3), dtype=t.float64, device=self.device)
        # the owned positions LIVE in the head of the local array (no copy per step)
        self.d_pos_owned = self.d_pos_loc[:self.n_owned]
        self.d_pos_owned.copy_(self._pos0)
        self.d_fp_halo = fp_halo if fp_halo is not None else \
            t.zeros(max(n_halo, 1), dtype=t.float64, device=self.device)
        if out is not None:
            self.d_out = out
        o = self.n_owned
        self.recv_pos_l = self.d_pos_loc[o:o + n_from_left]
        self.recv_pos_r = self.d_pos_loc[o + n_from_left:]
        self.recv_fp_l = self.d_fp_halo[:n_from_left]
        self.recv_fp_r = self.d_fp_halo[n_from_left:n_halo]

    def pack_positions(self, dst_l=None, dst_r=None):
        """Gather (+ periodic shift) the boundary atoms' positions into `dst_*`
        (default: the private send buffers; PeerComm passes the neighbours'
        receive regions, so the pack kernel IS the send)."""
        dst_l = self.send_pos_l if dst_l is None else dst_l
        dst_r = self.send_pos_r if dst_r is None else dst_r
        p = self.d_pos_owned
        self._lib.pack_rows(p, self.idx_l, dst_l, self.shift_l)
        self._lib.pack_rows(p, self.idx_r, dst_r, self.shift_r)
        return dst_l, dst_r

    def pack_fprime(self, dst_l=None, dst_r=None):
        dst_l = self.send_fp_l if dst_l is None else dst_l
        dst_r = self.send_fp_r if dst_r is None else dst_r
        self._lib.pack_rows(self.d_fp, self.idx_l, dst_l)
        self._lib.pack_rows(self.d_fp, self.idx_r, dst_r)
        return dst_l, dst_r

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
        self.graph = None
        r = self.rank_state
        self.peer = None
        import os
        if device == 'cuda' and os.environ.get('TAB_DD_PEER', '1') != '0':
            try:
                self.peer = PeerComm(self.layout, r.n_owned, len(r.idx_l), len(r.idx_r),
                                     device)
            except Exception as exc:      # no peer mappings: NCCL point-to-point
                self.peer = None
                self.peer_error = f"{type(exc).__name__}: {exc}"
        # every rank must take the same path
        import torch
        flag = torch.tensor([1 if self.peer is not None else 0], device=device)
        self.comm.dist.all_reduce(flag, op=self.comm.dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            self.peer = None
        if self.peer is not None:
            pc = self.peer
            r.set_halo_counts(pc.n_from_l, pc.n_from_r, pos_loc=pc.pos_loc,
                              fp_halo=pc.fp_halo, out=pc.red)
        else:
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
        how = ("pack kernels store into the ring neighbours' symmetric memory over "
               "NVLink + device barrier (positions, F'), one-shot peer all-reduce of 10 "
               "doubles") if self.peer is not None else \
              "2 NCCL ring exchanges (positions, F') + one 10-double all-reduce per step"
        return (f"1-D slabs along x, {lay.world} ranks x {self.n_local} owned atoms "
                f"(this rank), halo = rc = {lay.rc} A, {how}; "
                f"{'CUDA-graph step; ' if self.graph is not None else ''}"
                f"{self.scaling} scaling")

    def _exchange_positions(self):
        r = self.rank_state
        if self.peer is not None:
            r.pack_positions(self.peer.dst_pos_l, self.peer.dst_pos_r)
            self.peer.fence(0)
            return
        s_l, s_r = r.pack_positions()
        self.comm.exchange(s_l, s_r, r.recv_pos_l, r.recv_pos_r)

    def _exchange_fprime(self):
        r = self.rank_state
        if self.peer is not None:
            r.pack_fprime(self.peer.dst_fp_l, self.peer.dst_fp_r)
            self.peer.fence(1)
            return
        s_l, s_r = r.pack_fprime()
        self.comm.exchange(s_l, s_r, r.recv_fp_l, r.recv_fp_r)

    def _reduce(self):
        r = self.rank_state
        if self.peer is not None:
            self.peer.allreduce_sum()
        else:
            self.comm.allreduce_sum(r.d_out[:10])

    def _totals(self):
        return self.peer.red_out if self.peer is not None else self.rank_state.d_out

    def _step_body(self):
        r = self.rank_state
        self._exchange_positions()
        r.update()
        r.pass1()
        self._exchange_fprime()
        r.pass2()
        self._reduce()

    def enable_graph(self, warmup=3):
        """Capture the resident-list step (kernels, peer stores, barriers, reduction)
        in one CUDA graph.  Returns True when the capture worked."""
        import torch
        if self.peer is None and self.layout.world > 1:
            # NCCL point-to-point inside a stream capture is not robust (a capture that
            # fails on one rank leaves the ring waiting): the fallback path runs eagerly
            self.graph = None
            self.graph_error = "graph capture needs the peer-memory path"
            return False
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._step_body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_body()
            self.graph = g
        except Exception as exc:
            self.graph = None
            self.graph_error = f"{type(exc).__name__}: {exc}"
        # all ranks or none (a mixed ring would deadlock in the barriers)
        flag = torch.tensor([1 if self.graph is not None else 0], device='cuda')
        self.comm.dist.all_reduce(flag, op=self.comm.dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            self.graph = None
        return self.graph is not None

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()

    def step_e2e(self):
        """Host positions in, host forces / energy / virial out, lists rebuilt."""
        r = self.rank_state
        r.d_pos_owned.copy_(r.h_pos, non_blocking=True)
        self._exchange_positions()
        r.build()
        r.pass1()
        self._exchange_fprime()
        r.pass2()
        self._reduce()
        r.h_f.copy_(r.d_f, non_blocking=True)
        r.h_out.copy_(self._totals(), non_blocking=True)
        r.torch.cuda.synchronize()

    def results(self):
        r = self.rank_state
        out = self._totals().cpu().numpy()
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
