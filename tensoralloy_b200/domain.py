"""
Spatial (1-D slab) decomposition of the EAM force step over the GPUs of one box.

The reference has no spatial decomposition; this is the new
multi-GPU path of SURVEY.md 8(e):

  rank r owns the atoms with  lo_r <= x < hi_r  (slabs along x, periodic ring);
  per step
    1. pack the positions of the atoms within rc + skin of the two slab faces, exchange
       them with the two ring neighbours (halo = rc + skin; atoms crossing the periodic
       boundary are shifted by -/+ Lx by the sender)
    2. refresh the local lists over owned + halo atoms (tab_nbr_update, which also
       tracks the largest displacement since the build; y and z stay periodic inside
       the library)
    3. pass 1 on the owned atoms: rho_i, F(rho_i), F'(rho_i)
    4. exchange F' of the same boundary atoms (8 B per halo atom)
    5. pass 2 on the owned atoms: forces, energies, partial virial
    6. all-reduce of [E, virial(9)] (sum) and of the largest displacement (max)
  and, when an atom has moved (or may move in the next step) more than skin / 2,
    0. REBUILD: wrap the owned atoms, MIGRATE those that left the slab to the ring
       neighbour that now owns them (positions + whatever per-atom state the caller
       carries), recompute the send sets, exchange, tab_nbr_build_dd.
  With FULL neighbour lists an owned atom never needs a ghost's force, only its
  position and F'(rho): no reverse communication.  Lists carry a skin and the pair
  kernels mask r >= rc, so every step equals the evaluation on fresh lists.

`PeerComm` is the NVLink path: every rank's local position / F' arrays live in
torch symmetric memory, the pack kernels of a rank store STRAIGHT into its ring
neighbours' receive regions over NVLink (pack + send are one kernel, no staging
buffer, no NCCL call) and a device-side barrier orders the steps; the reduction is a
one-shot all-reduce over the same peer mappings.  `DistComm` is
the fallback over torch.distributed point-to-point (NCCL on GPUs, gloo in the CPU
tests) and carries the migration (variable-size messages, rebuild steps only).  The
resident-list step is captured in a CUDA graph (`enable_graph`, re-captured after a
rebuild), so a step is ONE launch per rank.  `run_loopback` runs every rank inside one
process to test the whole pipeline on a single GPU.
"""
import numpy as np


def _graph_exec_update(live, captured):
    """cudaGraphExecUpdate(exec of `live`, graph of `captured`) through the CUDA runtime that
    torch has loaded.  True when the executable graph now runs the newly captured work."""
    import ctypes
    try:
        rt = _graph_exec_update.rt
    except AttributeError:
        rt = None
        for name in ('libcudart.so.12', 'libcudart.so'):
            try:
                rt = ctypes.CDLL(name)
                break
            except OSError:
                continue
        _graph_exec_update.rt = rt
    if rt is None:
        return False
    try:
        info = (ctypes.c_ubyte * 64)()         # cudaGraphExecUpdateResultInfo (24 bytes)
        rt.cudaGraphExecUpdate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        rt.cudaGraphExecUpdate.restype = ctypes.c_int
        err = rt.cudaGraphExecUpdate(ctypes.c_void_p(int(live.raw_cuda_graph_exec())),
                                     ctypes.c_void_p(int(captured.raw_cuda_graph())),
                                     ctypes.cast(info, ctypes.c_void_p))
        if err != 0:
            rt.cudaGetLastError()               # clear the sticky error of a refused update
            return False
        return True
    except Exception:
        return False


class SlabLayout:
    """Pure geometry: which atoms a rank owns and which it sends where."""

    def __init__(self, lx, world, rank, rc, skin=0.0):
        self.lx = float(lx)
        self.world = int(world)
        self.rank = int(rank)
        self.rc = float(rc)
        self.skin = float(skin)
        self.reach = self.rc + self.skin        # radius of the lists = width of the halo
        self.width = self.lx / self.world
        if self.world > 1 and self.width < 2.0 * self.reach:
            raise ValueError(
                f"slab width {self.width:.3f} < 2 (rc + skin): the halo would span more "
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
        """Atoms within rc + skin of the low / high face (x_owned already wrapped)."""
        return x_owned < self.lo + self.reach, x_owned >= self.hi - self.reach

    def frame(self, ly, lz, pad=0.5):
        """Binning frame of the local system handed to tab_nbr_build_dd."""
        pad = pad + self.skin
        cell = np.diag([self.width + 2 * self.reach + 2 * pad, ly, lz])
        origin = np.array([self.lo - self.reach - pad, 0.0, 0.0])
        return cell, origin, [0, 1, 1]

    def owner_of(self, x_wrapped):
        """Rank that owns every wrapped x (numpy or torch)."""
        r = (x_wrapped / self.width).floor() if hasattr(x_wrapped, 'floor') else \
            np.floor(x_wrapped / self.width)
        return r.clip(0, self.world - 1) if not hasattr(r, 'clamp') else \
            r.clamp(0, self.world - 1)


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
        # empty messages are skipped on both sides (the receiver sized its buffer from the
        # sender's count)
        ops = [op for op in ops if op.tensor.numel() > 0]
        if not ops:
            return
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

    def allreduce_max(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)

    def migrate(self, state):
        """Rebuild-time re-assignment of owners.  `state`: [n, C] float64, columns 0..2 =
        positions of the atoms this rank owned so far (any C - 3 further per-atom columns
        travel along: velocities, ids, ...).  x is wrapped into [0, Lx); atoms whose x left
        the slab go to the ring neighbour that owns them now (an atom moves less than half a
        skin between two rebuilds, so never further).  Returns the new state:
        [kept | received from the left | received from the right]."""
        import torch
        lay = self.layout
        state = state.clone()
        state[:, 0] = torch.remainder(state[:, 0], lay.lx)
        # a coordinate that rounds to exactly Lx belongs to the first slab
        state[:, 0] = torch.where(state[:, 0] >= lay.lx, torch.zeros_like(state[:, 0]),
                                  state[:, 0])
        if lay.world == 1:
            return state
        owner = lay.owner_of(state[:, 0]).to(torch.int64)
        keep = owner == lay.rank
        to_l = owner == lay.left
        to_r = owner == lay.right
        if lay.world == 2:
            # left == right: everything that leaves goes to the one other rank ("to the
            # right"); the receive side mirrors it
            to_r = ~keep
            to_l = torch.zeros_like(keep)
        if not bool((keep | to_l | to_r).all()):
            raise RuntimeError("an atom moved further than to the adjacent slab between two "
                               "list rebuilds")
        s_l, s_r = state[to_l].contiguous(), state[to_r].contiguous()
        n_l, n_r = self.exchange_counts(int(s_l.shape[0]), int(s_r.shape[0]), state.device)
        c = state.shape[1]
        r_l = torch.empty((n_l, c), dtype=state.dtype, device=state.device)
        r_r = torch.empty((n_r, c), dtype=state.dtype, device=state.device)
        self.exchange(s_l, s_r, r_l, r_r)
        return torch.cat([state[keep], r_l, r_r], dim=0).contiguous()


class PeerComm:
    """Halo exchange by direct stores into the neighbours' memory (NVLink / NVSwitch
    peer mappings from torch symmetric memory).

    One flat float64 symmetric buffer per rank, same capacity everywhere:
        [ positions of owned | from-left | from-right atoms  (3 doubles each) ]
        [ F' of from-left | from-right atoms ]
    and a 16-double symmetric buffer per rank for [E, virial(9), max displacement].
    `dst_*` are views of the NEIGHBOURS' buffers: the region of the left neighbour that
    holds what it receives from its right (= me), and vice versa.  The buffers are
    allocated once (`slack` above the first layout); `configure` recomputes the views after
    every rebuild, when the owned / halo counts of the ranks have changed.
    """
    MIG_CAP = 16384      # atoms that may leave through one face between two rebuilds
    MIG_COLS = 7         # position, velocity, id

    def __init__(self, layout, n_owned, n_send_left, n_send_right, device, slack=1.12):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.torch, self.dist, self.layout = torch, dist, layout
        self.device = device
        world = layout.world
        table = self._gather_counts(n_owned, n_send_left, n_send_right)
        owned = [t[0] for t in table]
        from_l = [table[(r - 1) % world][2] for r in range(world)]
        from_r = [table[(r + 1) % world][1] for r in range(world)]
        self.rows_cap = int(slack * max(owned[r] + from_l[r] + from_r[r]
                                        for r in range(world))) + 64
        self.halo_cap = int(slack * max(from_l[r] + from_r[r] for r in range(world))) + 64
        self.off_fp = 3 * self.rows_cap
        total = self.off_fp + self.halo_cap
        group = dist.group.WORLD
        self.buf = symm.empty(total, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        # reduction of [E, virial, max disp]: slot r of every rank's `red_all` is written by r
        self.red_all = symm.empty(16 * world, dtype=torch.float64, device=device)
        self.red_all.zero_()
        self.red_hdl = symm.rendezvous(self.red_all, group)
        self.red_ptrs = torch.tensor([int(p) for p in self.red_hdl.buffer_ptrs],
                                     dtype=torch.int64, device=device)
        self.red = torch.zeros(16, dtype=torch.float64, device=device)      # my partial sums
        self.red_out = torch.zeros(16, dtype=torch.float64, device=device)
        # rebuild-time exchanges through the same peer mappings (no NCCL call, one read-back
        # each): the migration mailbox [2 directions][1 + MIG_CAP * MIG_COLS] (count, then
        # rows) and the table of every rank's [n_owned, n_send_left, n_send_right]
        self.mail = symm.empty(2 * (1 + self.MIG_CAP * self.MIG_COLS), dtype=torch.float64,
                               device=device)
        self.mail.zero_()
        self.mail_hdl = symm.rendezvous(self.mail, group)
        self.counts = symm.empty(4 * world, dtype=torch.float64, device=device)
        self.counts.zero_()
        self.counts_hdl = symm.rendezvous(self.counts, group)
        self.counts_ptrs = torch.tensor([int(p) for p in self.counts_hdl.buffer_ptrs],
                                        dtype=torch.int64, device=device)
        L, R = layout.left, layout.right
        box = 1 + self.MIG_CAP * self.MIG_COLS
        # what I send to the left lands in the left neighbour's "from the right" box, ...
        self.mail_to_l = self.mail_hdl.get_buffer(L, (box,), torch.float64, box)
        self.mail_to_r = self.mail_hdl.get_buffer(R, (box,), torch.float64, 0)
        # the migration's two alternating row buffers, its scratch and its counters: allocated
        # here, not inside the first rebuilds (an allocation synchronises the device, and under
        # a graph's private pool it was measured at several milliseconds)
        cap = int(slack * max(owned)) + 2 * self.MIG_CAP + 64
        self._mig_bufs = [torch.empty((cap, self.MIG_COLS), dtype=torch.float64, device=device)
                          for _ in range(2)]
        self._mig_flip = 0
        self._mig_work = torch.empty(2 * (3 * ((cap + 255) // 256) + 8), dtype=torch.int32,
                                     device=device)
        self._mig_counts = torch.zeros(8, dtype=torch.int32, device=device)
        self._apply(table, n_send_left, n_send_right)
        torch.cuda.synchronize()
        dist.barrier()

    def _gather_counts(self, n_owned, n_send_left, n_send_right):
        torch, dist = self.torch, self.dist
        world = self.layout.world
        mine = torch.tensor([n_owned, n_send_left, n_send_right], dtype=torch.int64,
                            device=self.device)
        table = [torch.zeros(3, dtype=torch.int64, device=self.device) for _ in range(world)]
        dist.all_gather(table, mine)
        return [t.tolist() for t in table]

    def configure(self, n_owned, n_send_left, n_send_right):
        """New owned / send counts (after a migration): recompute every view.  The counts
        travel through the symmetric count table: one kernel stores mine into every rank's
        table, a device barrier, one read-back."""
        from tensoralloy_b200 import _lib
        torch = self.torch
        world, rank = self.layout.world, self.layout.rank
        mine = torch.tensor([n_owned, n_send_left, n_send_right, 0.0], dtype=torch.float64,
                            device=self.device)
        _lib.peer_put(mine, self.counts_ptrs, rank)
        self.counts_hdl.barrier(channel=3)
        table = self.counts.view(world, 4)[:, :3].to(torch.int64).tolist()
        self._apply(table, n_send_left, n_send_right)

    def migrate(self, state):
        """`DistComm.migrate` over the peer mappings, on the device: `tab_dd_partition` wraps x,
        classifies every row (keep / left / right neighbour), compacts the kept rows and stores
        the leaving rows with their count straight into the ring neighbours' mailboxes (three
        launches, stable order); one device barrier; ONE read-back of [kept, left, right, lost,
        overflow] and the two received counts.  The received rows are appended in a fixed order
        (from the left, then from the right; each in the sender's order).  Returns a view of
        one of two alternating buffers."""
        from tensoralloy_b200 import _lib
        torch = self.torch
        lay = self.layout
        n, c = int(state.shape[0]), self.MIG_COLS
        state = state.contiguous()
        cap = n + 2 * self.MIG_CAP
        self._mig_flip = 1 - self._mig_flip
        bufs = self._mig_bufs
        keep = bufs[self._mig_flip]
        if keep.shape[0] < cap or keep.data_ptr() == state.data_ptr():
            keep = bufs[self._mig_flip] = torch.empty((cap + cap // 8, c), dtype=torch.float64,
                                                      device=self.device)
        nwork = 3 * ((n + 255) // 256) + 8
        if self._mig_work.numel() < nwork:
            self._mig_work = torch.empty(2 * nwork, dtype=torch.int32, device=self.device)
        _lib.dd_partition(state, lay.lx, lay.width, lay.world, lay.rank, keep, self.mail_to_l,
                          self.mail_to_r, self.MIG_CAP, self._mig_counts, self._mig_work)
        self.mail_hdl.barrier(channel=4)
        box = 1 + self.MIG_CAP * c
        got = torch.cat((self._mig_counts[:5].to(torch.float64), self.mail[0:1],
                         self.mail[box:box + 1])).tolist()                    # the read-back
        n_keep, n_l, n_r, n_lost, overflow, a, b = (int(v) for v in got)
        if n_lost:
            raise RuntimeError("an atom moved further than to the adjacent slab between two "
                               "list rebuilds")
        if overflow:
            raise RuntimeError(f"{max(n_l, n_r)} atoms leave through one face: more than the "
                               f"migration mailbox holds ({self.MIG_CAP})")
        if a:
            keep[n_keep:n_keep + a] = self.mail[1:1 + a * c].view(a, c)
        if b:
            keep[n_keep + a:n_keep + a + b] = self.mail[box + 1:box + 1 + b * c].view(b, c)
        # (a mailbox is not overwritten before its owner has copied it out: the peers' next
        # stores into it come after the device barrier of `configure`, which this rank enters
        # after the copies above, in stream order)
        return keep[:n_keep + a + b]

    def configure_device(self, n_owned, d_send_counts):
        """`configure` with the send counts still on the device (int32 [2] written by
        `tab_dd_send_sets`): they go into the count table without a detour through the host;
        the read-back of the table is the only synchronisation.  Returns (n_send_left,
        n_send_right) of this rank."""
        from tensoralloy_b200 import _lib
        torch = self.torch
        world, rank = self.layout.world, self.layout.rank
        mine = torch.empty(4, dtype=torch.float64, device=self.device)
        mine[0] = float(n_owned)
        mine[1:3] = d_send_counts[:2]
        mine[3] = 0.0
        _lib.peer_put(mine, self.counts_ptrs, rank)
        self.counts_hdl.barrier(channel=3)
        table = self.counts.view(world, 4)[:, :3].to(torch.int64).tolist()
        n_l, n_r = table[rank][1], table[rank][2]
        self._apply(table, n_l, n_r)
        return n_l, n_r

    def _apply(self, table, n_send_left, n_send_right):
        torch = self.torch
        lay = self.layout
        world, rank = lay.world, lay.rank
        owned = [t[0] for t in table]
        from_l = [table[(r - 1) % world][2] for r in range(world)]
        from_r = [table[(r + 1) % world][1] for r in range(world)]
        if max(owned[r] + from_l[r] + from_r[r] for r in range(world)) > self.rows_cap or \
                max(from_l[r] + from_r[r] for r in range(world)) > self.halo_cap:
            raise RuntimeError("slab populations outgrew the symmetric buffers "
                               f"(capacity {self.rows_cap} rows); construct the domain with "
                               "a larger slack")
        self.n_from_l, self.n_from_r = from_l[rank], from_r[rank]
        L, R = lay.left, lay.right
        f64 = torch.float64
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
        # (no barrier here: a rebuild follows a completed step, whose last device barrier
        # every rank has passed -- nobody reads the old regions any more; the all_gather of the
        # counts orders the ranks before anyone stores into the new ones)

    def fence(self, channel):
        """All ranks have issued (and completed) the stores before this point."""
        self.hdl.barrier(channel=channel)

    def allreduce(self):
        """self.red over the ranks -> self.red_out: entries 0..9 ([E, virial]) summed,
        entry 10 (largest displacement) reduced with max.  One kernel stores my 16 doubles
        into my slot of EVERY rank's buffer over NVLink, a device barrier, one kernel
        reduces the slots in rank order (deterministic, identical on all ranks).  (torch's
        one-shot symmetric all-reduce has no float64 kernel.)"""
        from tensoralloy_b200 import _lib
        _lib.peer_put(self.red, self.red_ptrs, self.layout.rank)
        self.red_hdl.barrier(channel=2)
        _lib.sum_slots(self.red_all, self.layout.world, self.red_out, n_sum=10)
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
        self.nbr = _lib.NeighborList()
        self.nbr.set_skin(layout.skin)
        self.n_from_l = self.n_from_r = 0
        self.d_pos_loc = None
        self.set_owned(pos_owned)

    def set_owned(self, pos_owned):
        """(Re)define the owned atoms: numpy or device tensor [n, 3], wrapped x."""
        self.set_owned_begin(pos_owned)
        n_l, n_r = self.d_send_counts.tolist()          # one read-back for both counts
        self.set_owned_finish(n_l, n_r)

    def set_owned_begin(self, pos_owned):
        """Owned positions + the send sets (atoms within rc + skin of the faces) as ONE library
        call, `tab_dd_send_sets`: both index lists and their lengths stay on the device
        (`d_send_counts`); `set_owned_finish` takes the lengths once the host knows them."""
        torch = self.torch
        device = self.device
        if not torch.is_tensor(pos_owned):
            pos_owned = torch.from_numpy(np.ascontiguousarray(pos_owned))
        self._pos0 = pos_owned.to(device).contiguous()
        n = self.n_owned = int(self._pos0.shape[0])
        if getattr(self, '_idx_buf', None) is None or self._idx_buf.shape[1] < n:
            self._idx_buf = torch.empty((2, n + n // 8 + 64), dtype=torch.int64, device=device)
            self._send_work = torch.empty(2 * ((n + n // 8 + 64 + 255) // 256) + 8,
                                          dtype=torch.int32, device=device)
            self.d_send_counts = torch.zeros(2, dtype=torch.int32, device=device)
        self._lib.dd_send_sets(self._pos0, self.lay.lo + self.lay.reach,
                               self.lay.hi - self.lay.reach, self._idx_buf[0],
                               self._idx_buf[1], self.d_send_counts, self._send_work)

    def set_owned_finish(self, n_l, n_r):
        torch = self.torch
        device = self.device
        self.idx_l = self._idx_buf[0, :n_l]
        self.idx_r = self._idx_buf[1, :n_r]
        self.shift_l = [self.lay.shift_to_left, 0.0, 0.0]
        self.shift_r = [self.lay.shift_to_right, 0.0, 0.0]
        f64 = dict(dtype=torch.float64, device=device)
        self.send_pos_l = torch.empty((n_l, 3), **f64)
        self.send_pos_r = torch.empty((n_r, 3), **f64)
        self.send_fp_l = torch.empty(n_l, **f64)
        self.send_fp_r = torch.empty(n_r, **f64)
        self.d_fp = torch.zeros(self.n_owned, **f64)
        self.d_out = torch.zeros(16, **f64)
        self.d_f = torch.zeros((self.n_owned, 3), **f64)
        self.d_ea = None
        if getattr(self, 'h_out', None) is None:
            # once per rank: pinning host memory synchronises the device and can take
            # milliseconds (measured: a 10 ms outlier inside a rebuild)
            self.h_out = torch.zeros(16, dtype=torch.float64).pin_memory() \
                if device != 'cpu' else torch.zeros(16, dtype=torch.float64)
        self.h_f = None
        self.h_pos = None

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

    def pass2(self, eatom=None):
        self.model.pass2(self.nbr, self.precision,
                         fprime_halo=self.d_fp_halo if self.n_from_l + self.n_from_r
                         else None,
                         energy=self.d_out[0:1], eatom=eatom, forces=self.d_f,
                         virial=self.d_out[1:10])

    def displacement(self):
        """Largest displacement since the build -> d_out[10] (lists with a skin)."""
        if self.lay.skin > 0.0:
            self.nbr.displacement_to(self.d_out[10:11])


class SlabDomain:
    """One rank of the distributed MD force step (used by bench.py).

    `state` = [n_owned, 7] float64 on the device: position, velocity (Angstrom per step; the
    benchmark's stand-in for an integrator), global atom id.  `md_step` advances the atoms,
    refreshes or rebuilds (with migration) the lists, evaluates E / F / virial."""
    md_valid = True

    def __init__(self, model, cells, a, rc, sigma, seed, world, rank,
                 scaling='strong', precision=0, device='cuda', skin=0.0, vel_seed=None):
        import torch
        from tensoralloy_b200.atoms import fcc_positions
        self.torch = torch
        gx = cells * world if scaling == 'weak' else cells
        pos, cell = fcc_positions(a, gx, cells, cells)
        rng = np.random.default_rng(seed)
        pos = pos + rng.normal(scale=sigma, size=pos.shape)
        lx, ly, lz = cell[0, 0], cell[1, 1], cell[2, 2]
        pos[:, 0] = np.mod(pos[:, 0], lx)
        self.cell = cell
        self.n_total = len(pos)
        self.skin = float(skin)
        self.model = model
        self.precision = precision
        self.device = device
        self.ly, self.lz = ly, lz
        self.layout = SlabLayout(lx, world, rank, rc, skin)
        mask = self.layout.owned_mask(pos[:, 0])
        ids = np.flatnonzero(mask)
        # velocity field of the MD cycle (bench.py): the same Gaussian field on every world
        # size (drawn for all atoms, indexed by id), the fastest atom covers skin / 2 in
        # 9.5 steps
        vrng = np.random.default_rng(seed + 1 if vel_seed is None else vel_seed)
        vel = vrng.normal(size=pos.shape)
        vmax = np.linalg.norm(vel, axis=1).max()
        vel *= (0.5 * max(self.skin, 1e-3) / 9.5) / vmax
        self.vstep_max = 0.5 * max(self.skin, 1e-3) / 9.5
        state = np.concatenate([pos[mask], vel[mask], ids[:, None].astype(np.float64)], axis=1)
        del pos, vel
        self.state = torch.from_numpy(np.ascontiguousarray(state)).to(device)
        self.comm = DistComm(self.layout)
        self.scaling = scaling
        self.graph = None
        self._want_graph = False
        self.peer = None
        self.rebuilds = 0
        self.rebuild_profile = None      # dict: accumulates ms per phase of `rebuild`
        self._since_build = -1          # md steps since the lists were built (-1: fresh lists)
        self._disp_ring = None
        self.d_move = torch.zeros(1, dtype=torch.float64, device=device)   # 1: atoms advance
        self._moving = False
        self.rank_state = SlabRank(model, self.layout, self.state[:, 0:3], ly, lz, precision,
                                   device)
        r = self.rank_state
        import os
        if device == 'cuda' and os.environ.get('TAB_DD_PEER', '1') != '0':
            try:
                self.peer = PeerComm(self.layout, r.n_owned, len(r.idx_l), len(r.idx_r),
                                     device)
            except Exception as exc:      # no peer mappings: NCCL point-to-point
                self.peer = None
                self.peer_error = f"{type(exc).__name__}: {exc}"
        # every rank must take the same path
        flag = torch.tensor([1 if self.peer is not None else 0], device=device)
        self.comm.dist.all_reduce(flag, op=self.comm.dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            self.peer = None
        self._attach(first=True)

    # -- (re)attachment of the rank state to the communication buffers -----------
    def _attach(self, first=False, lap=None, configured=False):
        lap = lap or (lambda name: None)
        r = self.rank_state
        if self.peer is not None:
            pc = self.peer
            if not first and not configured:
                pc.configure(r.n_owned, len(r.idx_l), len(r.idx_r))
            r.set_halo_counts(pc.n_from_l, pc.n_from_r, pos_loc=pc.pos_loc,
                              fp_halo=pc.fp_halo, out=pc.red)
        else:
            n_l, n_r = self.comm.exchange_counts(len(r.idx_l), len(r.idx_r), self.device)
            r.set_halo_counts(n_l, n_r)
        self.d_vel = self.state[:, 3:6].contiguous()
        lap('layout')
        self._exchange_positions()
        lap('halo_exchange')
        r.build()
        lap('list_build')
        self.n_local = r.n_owned
        self.nij_local = r.nbr.sizes()[0]
        self.h2d_bytes = r.n_owned * 24
        self.d2h_bytes = r.n_owned * 24 + 80

    def set_precision(self, name):
        self.precision = self.prec_id[name]
        self.rank_state.precision = self.precision
        # the captured kernels depend on the precision: the next steps run eagerly (first use
        # of a precision allocates inside the library) and `md_step` captures again
        self.graph = None

    def describe(self):
        lay = self.layout
        how = ("pack kernels store into the ring neighbours' symmetric memory over "
               "NVLink + device barrier (positions, F'), one-shot peer all-reduce of "
               "[E, virial, max displacement]") if self.peer is not None else \
              "2 NCCL ring exchanges (positions, F') + one all-reduce per step"
        return (f"1-D slabs along x, {lay.world} ranks x {self.n_local} owned atoms "
                f"(this rank), halo = rc + skin = {lay.reach} A, {how}; "
                f"{'CUDA-graph step; ' if self.graph is not None else ''}"
                f"atoms migrate between slabs at every rebuild; {self.scaling} scaling")

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
            self.peer.allreduce()
        else:
            self.comm.allreduce_sum(r.d_out[:10])
            if self.skin > 0.0:
                self.comm.allreduce_max(r.d_out[10:11])

    def _totals(self):
        return self.peer.red_out if self.peer is not None else self.rank_state.d_out

    def _step_body(self, eatom=None):
        r = self.rank_state
        # the integrator's stand-in: R += move * v (move = 0 for the resident-step timing)
        r.d_pos_owned.addcmul_(self.d_vel, self.d_move)
        self._exchange_positions()
        r.update()
        r.displacement()
        r.pass1()
        self._exchange_fprime()
        r.pass2(eatom)
        self._reduce()

    def enable_graph(self, warmup=2):
        """Capture the resident-list step (kernels, peer stores, barriers, reduction)
        in one CUDA graph.  Returns True when the capture worked.  The graph is
        re-captured after every rebuild (the owned / halo counts change): the re-captures
        share the first capture's memory pool and side stream and skip the warm-up."""
        import torch
        first = not self._want_graph or getattr(self, '_graph_pool', None) is None
        self._want_graph = True
        if self.peer is None and self.layout.world > 1:
            # NCCL point-to-point inside a stream capture is not robust (a capture that
            # fails on one rank leaves the ring waiting): the fallback path runs eagerly
            self.graph = None
            self.graph_error = "graph capture needs the peer-memory path"
            return False
        try:
            if first:
                # warm-up launches must not move the atoms
                self._set_moving(False)
                self._graph_stream = torch.cuda.Stream()
                side = self._graph_stream
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(warmup):
                        self._step_body()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                self._graph_pool = torch.cuda.graph_pool_handle()
            # low-level capture: the torch.cuda.graph context manager runs gc.collect() and
            # empty_cache() on entry (5 ms per capture, measured) -- a re-capture happens at
            # every list rebuild.  keep_graph: the cudaGraph_t stays available, so that a
            # re-capture can UPDATE the parameters of the existing executable graph
            # (cudaGraphExecUpdate: same kernels, new counts / pointers) instead of
            # instantiating a new one (3 ms -> 0.1 ms, profiles/r02g).
            g = torch.cuda.CUDAGraph(keep_graph=True)
            side = self._graph_stream
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                # thread_local: CUDA calls of other threads (the NCCL watchdog polls events)
                # must not invalidate the capture
                g.capture_begin(pool=self._graph_pool, capture_error_mode="thread_local")
                try:
                    self._step_body()
                finally:
                    g.capture_end()
            torch.cuda.current_stream().wait_stream(side)
            # the capture itself does not run the kernels: nothing has moved
            live = getattr(self, '_graph_exec', None)
            if live is not None and _graph_exec_update(live, g):
                self.graph = live           # the old executable, new parameters
                self._graph_src = g         # (keeps the captured cudaGraph_t alive)
            else:
                g.instantiate()
                self.graph = self._graph_exec = g
        except Exception as exc:
            self.graph = None
            self.graph_error = f"{type(exc).__name__}: {exc}"
        if first:
            # one answer for the whole job (bench.py reports it).  Re-captures do not need the
            # agreement: a replayed graph and an eager step issue the same kernels, stores and
            # device barriers, so ranks may differ without waiting on each other
            flag = torch.tensor([1 if self.graph is not None else 0], device='cuda')
            self.comm.dist.all_reduce(flag, op=self.comm.dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                self.graph = None
        if self.graph is None:
            self._want_graph = False        # do not try again after every rebuild
        return self.graph is not None

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()

    def _set_moving(self, on):
        if on != self._moving:
            self.d_move.fill_(1.0 if on else 0.0)
            self._moving = on

    # -- MD cycle ----------------------------------------------------------------
    def rebuild(self):
        """Migrate, recompute the send sets, exchange, rebuild the lists (collective)."""
        import time
        r = self.rank_state
        torch = self.torch
        prof = self.rebuild_profile
        tick = [time.perf_counter()]

        def lap(name):
            if prof is not None:
                if self.device != 'cpu':
                    torch.cuda.synchronize()
                now = time.perf_counter()
                prof[name] = prof.get(name, 0.0) + (now - tick[0]) * 1e3
                tick[0] = now

        self.state[:, 0:3] = r.d_pos_owned
        self.state = self.peer.migrate(self.state) if self.peer is not None else \
            self.comm.migrate(self.state)
        lap('migrate')
        configured = False
        if self.peer is not None:
            # send sets and the exchange of the new counts without a host detour: one read-back
            r.set_owned_begin(self.state[:, 0:3])
            n_l, n_r = self.peer.configure_device(r.n_owned, r.d_send_counts)
            r.set_owned_finish(n_l, n_r)
            configured = True
        else:
            r.set_owned(self.state[:, 0:3])
        lap('send_sets')
        self.graph = None               # (the executable graph survives in _graph_exec)
        self._attach(lap=lap, configured=configured)
        self._since_build = -1          # fresh lists, no step on them yet
        self.rebuilds += 1
        # the step that follows runs eagerly (with the new sizes: every library buffer that has
        # to grow does so outside a capture); `md_step` re-captures after it

    def md_step(self):
        """One MD step with moving atoms.  The decision to rebuild is taken BEFORE the lists
        could become invalid and WITHOUT waiting for the device: the largest displacement since
        the build (max over the ranks, part of every step's reduction) is read back
        asynchronously; step s after a build knows the reading of step s - 2 (long finished) and
        rebuilds first when that reading plus two steps at the fastest atom's speed could exceed
        skin / 2.  Every rank sees the same readings, so the decision is collective without a
        message; the host runs one step ahead of the device instead of stalling it at every
        step (at 8 GPUs the stall was > 10 % of a 0.28 ms step)."""
        torch = self.torch
        s = self._since_build + 1          # index of this step since the lists were built
        if self.skin <= 0.0:
            need = True
        else:
            if s <= 1:
                bound = (s + 1) * self.vstep_max
            else:
                ev, buf = self._disp_ring[s % 2]          # written by step s - 2
                ev.synchronize()
                bound = float(buf[0]) + 2.0 * self.vstep_max
            need = not (2.0 * bound <= self.skin)
        if need:
            self.rebuild()
            s = 0
        self._since_build = s
        self._set_moving(True)
        self.step()
        if self.graph is None and self._want_graph and self.skin > 0.0:
            import time
            if self.rebuild_profile is not None:
                torch.cuda.synchronize()            # (the eager step is not part of it)
            t0 = time.perf_counter()
            self.enable_graph()
            if self.rebuild_profile is not None:
                torch.cuda.synchronize()
                self.rebuild_profile['graph_capture'] = \
                    self.rebuild_profile.get('graph_capture', 0.0) + \
                    (time.perf_counter() - t0) * 1e3
        if self.skin > 0.0:
            if self._disp_ring is None:
                self._disp_ring = [(torch.cuda.Event(),
                                    torch.zeros(1, dtype=torch.float64).pin_memory())
                                   for _ in range(2)]
            ev, buf = self._disp_ring[s % 2]
            buf.copy_(self._totals()[10:11], non_blocking=True)
            ev.record()

    def resident_step(self):
        self._set_moving(False)
        self.step()

    def step_e2e(self):
        """Host positions in, host forces / energy / virial out, lists rebuilt."""
        r = self.rank_state
        torch = self.torch
        if r.h_pos is None or r.h_pos.shape[0] != r.n_owned:
            r.h_pos = r.d_pos_owned.cpu().pin_memory()
            r.h_f = torch.zeros((r.n_owned, 3), dtype=torch.float64).pin_memory()
        r.d_pos_owned.copy_(r.h_pos, non_blocking=True)
        self._exchange_positions()
        r.build()
        self._since_build = -1          # fresh lists at the current positions
        r.pass1()
        self._exchange_fprime()
        r.pass2()
        self._reduce()
        r.h_f.copy_(r.d_f, non_blocking=True)
        r.h_out.copy_(self._totals(), non_blocking=True)
        torch.cuda.synchronize()

    def results(self):
        r = self.rank_state
        out = self._totals().cpu().numpy()
        return out[0], r.d_f.cpu().numpy(), out[1:10].reshape(3, 3)

    def gather_global(self):
        """(positions [N,3], forces [N,3]) of ALL atoms in id order on every rank (checks)."""
        torch, dist = self.torch, self.comm.dist
        r = self.rank_state
        world = self.layout.world
        mine = torch.cat([self.state[:, 6:7], r.d_pos_owned, r.d_f], dim=1).contiguous()
        counts = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([mine.shape[0]], dtype=torch.int64,
                                             device=self.device))
        cap = max(int(c.item()) for c in counts)
        pad = torch.zeros((cap, 7), dtype=torch.float64, device=self.device)
        pad[:mine.shape[0]] = mine
        parts = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        rows = torch.cat([p[:int(c.item())] for p, c in zip(parts, counts)], dim=0).cpu().numpy()
        order = np.argsort(rows[:, 0].astype(np.int64))
        rows = rows[order]
        assert len(rows) == self.n_total and \
            np.array_equal(rows[:, 0].astype(np.int64), np.arange(self.n_total)), \
            "atoms were lost or duplicated by the migration"
        return rows[:, 1:4], rows[:, 4:7]

    def check(self, n_sample, dist=None, sampled=None):
        """Parity evidence of the decomposed run (bench.py `check`): the oracle on the 2 rc
        environments of sampled atoms (every rank checks a share of ITS atoms), the energy /
        force checksums, and reused-vs-fresh lists."""
        torch = self.torch
        r = self.rank_state
        d = self.comm.dist
        world = self.layout.world
        ea = torch.zeros(r.n_owned, dtype=torch.float64, device=self.device)
        graph, self.graph = self.graph, None
        self._set_moving(False)
        self._step_body(eatom=ea)
        torch.cuda.synchronize()
        tot = self._totals().cpu().numpy().copy()
        f_skin = r.d_f.clone()
        pos_all, f_all = self.gather_global()
        out = {}
        if sampled is not None:
            share = max(1, n_sample // world)
            rng = np.random.default_rng(611 + self.layout.rank)
            loc = np.sort(rng.choice(r.n_owned, size=min(share, r.n_owned), replace=False))
            ids = self.state[:, 6].cpu().numpy().astype(np.int64)[loc]
            t_loc = torch.from_numpy(loc).to(self.device)
            max_de, max_df = sampled(pos_all, self.cell, ids, ea[t_loc].cpu().numpy(),
                                     r.d_f[t_loc].cpu().numpy())
            errs = torch.tensor([max_de, max_df], dtype=torch.float64, device=self.device)
            d.all_reduce(errs, op=d.ReduceOp.MAX)
            n_s = torch.tensor([len(loc)], device=self.device)
            d.all_reduce(n_s)
            high = self.precision == 0
            tol_e, tol_f = (1e-10, 1e-8) if high else \
                (1e-5 * 4.45, 1e-5 * float(np.abs(f_all).max()))
            out.update({"n_sampled": int(n_s.item()), "max_dE_atom": float(errs[0].item()),
                        "max_dF": float(errs[1].item()), "tol_dE_atom": tol_e, "tol_dF": tol_f,
                        "ok": bool(errs[0].item() <= tol_e and errs[1].item() <= tol_f),
                        "against": "oracle (CPU restatement of the reference) on the 2 rc "
                                   "environment of every sampled atom (each rank samples its "
                                   "own atoms), positions of the last timed step"})
        f2 = (r.d_f * r.d_f).sum().reshape(1)
        d.all_reduce(f2)
        # the same positions on freshly built exact lists (the reference's per-call rebuild)
        skin_nbr = r.nbr
        r.nbr = self._lib_exact_nbr()
        try:
            r.build()
            r.pass1()
            self._exchange_fprime()
            r.pass2()
            self._reduce()
            torch.cuda.synchronize()
            tot_x = self._totals().cpu().numpy().copy()
            df = (f_skin - r.d_f).abs().max().reshape(1)
            d.all_reduce(df, op=d.ReduceOp.MAX)
        finally:
            r.nbr = skin_nbr
        n = self.n_total
        out.update({"energy": float(tot[0]), "f_l2": float(torch.sqrt(f2).item()),
                    "virial_trace": float(tot[1] + tot[5] + tot[9]),
                    "max_disp_since_build": float(tot[10]), "skin": self.skin,
                    "atoms_accounted_for": int(len(pos_all)),
                    "reused_vs_fresh_lists": {
                        "dE_per_atom": abs(float(tot[0]) - float(tot_x[0])) / n,
                        "max_dF": float(df.item()),
                        "max_dvirial_per_atom": float(np.abs(tot[1:10] - tot_x[1:10]).max()) / n}})
        self.graph = graph
        return out

    def _lib_exact_nbr(self):
        from tensoralloy_b200 import _lib
        if getattr(self, '_exact_nbr', None) is None:
            self._exact_nbr = _lib.NeighborList()
        return self._exact_nbr


def run_loopback(model, pos, cell, rc, world, precision=0, rebuild=True, device='cuda',
                 skin=0.0, moves=None):
    """Run every rank of a `world`-way slab decomposition inside ONE process on
    one GPU (test harness for the decomposition logic + the DD kernels).
    `moves`: optional list of displacement fields [N,3]; after the first evaluation the
    atoms are moved by each field in turn and re-evaluated on the REUSED lists
    (tab_nbr_update; lists with a skin).  Returns (E_total, forces[N,3] in input order,
    virial[3,3]) of the last evaluation."""
    import torch
    pos = np.array(pos, dtype=np.float64)
    lx, ly, lz = cell[0, 0], cell[1, 1], cell[2, 2]
    pos[:, 0] = np.mod(pos[:, 0], lx)
    ranks, owners = [], []
    for r in range(world):
        lay = SlabLayout(lx, world, r, rc, skin)
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

    def evaluate(build):
        exchange('pack_positions')
        for st in ranks:
            st.build() if build else st.update()
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

    out = evaluate(rebuild)
    for dR in (moves or []):
        for st, own in zip(ranks, owners):
            st.d_pos_owned.add_(torch.from_numpy(np.ascontiguousarray(dR[own])).to(device))
        out = evaluate(False)
    return out
